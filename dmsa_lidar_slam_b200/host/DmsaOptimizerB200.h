// DmsaOptimizerB200.h — header-only C++ adapter over the C-ABI (include/dmsa_b200.h).
//
// Keeps the reference's optimizer surface:
//     DmsaOptimizer<PointT>::optimizeSet(OptimizablePointSet<PointT>&, DmsaOptimSettings)      DmsaOptimizer.h:41-54
//     struct DmsaOptimSettings (same fields, same defaults)                                     DmsaOptimizer.h:25-39
// and is what DmsaSlam.h:52-53 would instantiate instead of DmsaOptimizer<PointStampId> /
// DmsaOptimizer<PointNormal>.
//
// Two build modes:
//   * default (this repository: no Eigen / PCL in the image): the adapter works on the POD "views" below, which carry
//     exactly the members of ContinuousTrajectory / MapManagement that the hot path reads and writes.
//   * -DDMSA_B200_WITH_REFERENCE_TYPES (inside the reference's catkin build): additionally defines
//     DmsaOptimizerB200<PointT> taking the real OptimizablePointSet<PointT>& and filling the views from the real objects.
//
// There is no CPU fallback: an unsupported point-set subclass or a CUDA failure throws std::runtime_error.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dmsa_b200.h"

namespace dmsa_b200 {

// DmsaOptimizer.h:25-39 — identical names, types and defaults
struct DmsaOptimSettings {
    int num_iter = 15;
    double epsilon = 1e-5;
    bool use_analytic_jacobi = false;
    double step_length_optim = 0.05;
    double max_step = 0.01;
    bool gauss_split = false;
    float grid_size_1_factor = 2.0;
    float grid_size_2_factor = 5.0;
    int min_num_points_per_set = 6;
    int min_num_gaussians = 30;
    float lambda_diag = 0.00001;
    bool use_centralization = true;
};

inline dmsa_b200_settings to_c(const DmsaOptimSettings& s) {
    dmsa_b200_settings c;
    c.num_iter = s.num_iter;
    c.epsilon = s.epsilon;
    c.use_analytic_jacobi = s.use_analytic_jacobi;
    c.step_length_optim = s.step_length_optim;
    c.max_step = s.max_step;
    c.gauss_split = s.gauss_split;
    c.grid_size_1_factor = s.grid_size_1_factor;
    c.grid_size_2_factor = s.grid_size_2_factor;
    c.min_num_points_per_set = s.min_num_points_per_set;
    c.min_num_gaussians = s.min_num_gaussians;
    c.lambda_diag = s.lambda_diag;
    c.use_centralization = s.use_centralization;
    return c;
}

// --- the members of ContinuousTrajectory the hot path touches (ContinuousTrajectory.h:24-72) ---------------------
struct ScanView {
    const dmsa_b200_point_stamp_id* points;  // PointCloudPlus::points.data()  (PointStampId is layout-identical)
    int64_t size;
    float gridSize;                          // PointCloudPlus::gridSize (PointCloudPlus.h:15-18)
};
struct TrajectoryView {
    double t_min = 0, t_max = 0, dt_res = 1e-3;  // arguments of initTraj (:301)
    // when > 0: the `horizon` member initTraj left behind (:309) with t_min = t0; passed through unchanged instead of
    // re-deriving it from t_max (an ulp there can move trajTime and flip a lower_bound of registerPcBuffer)
    double horizon = 0;
    int numControlPoses = 0;
    bool useImuErrorTerms = false;
    std::vector<ScanView> scans;                               // regPcBuffer, chronological (RingBuffer.h:31-37)
    const dmsa_b200_point_stamp_id* staticPoints = nullptr;    // tail of globalPoints with isStatic = 1 (:158-172)
    int64_t numStatic = 0;
    double* relOrientations = nullptr;   // controlPoses.relativePoses.Orientations.data()  (3 x n, in/out)
    double* relTranslations = nullptr;   // controlPoses.relativePoses.Translations.data()
    double* globOrientations = nullptr;  // controlPoses.globalPoses (out, may be null)
    double* globTranslations = nullptr;
    float* globalPointsXYZW = nullptr;   // out: N x 4 floats to scatter back into globalPoints[i].data (may be null)
    // IMU factor constants (only read when useImuErrorTerms): preintImuRots etc. (:36-43), row-major doubles
    const double* preintRot = nullptr;
    const double* preintPos = nullptr;
    const double* preintVel = nullptr;
    const double* covInv = nullptr;
    double balancingImu = 0.001f;
    const double* gravity = nullptr;     // Vector3d gravity (:34); nullptr = the initTraj default (0, 0, -9.805)
    // out (may be null): denseGlobalPoses (3 x n_total column-major doubles each, :29) and denseTformsLocal2Global as
    // n_total x 12 floats (rows 0..2 of each Matrix4f, :31) at the final parameters (DmsaOptimizer.h:149)
    double* denseOrientations = nullptr;
    double* denseTranslations = nullptr;
    float* denseTforms = nullptr;
};

// --- the members of MapManagement the hot path touches (MapManagement.h:20-70, KeyframeData.h:17-33) ----------------
struct KeyframeView {
    const dmsa_b200_point_normal* points;  // keyframeDataBuffer.at(k).pointCloudLocal->points.data()
    const int32_t* ringIds;                // keyframeDataBuffer.at(k).ringIds.data()
    int64_t size;
    float gridSize;
};
struct SubmapView {
    std::vector<KeyframeView> keyframes;
    // additional factors (MapManagement.h:39-54, 162-252; KeyframeData.h:23-31), n entries each, row-major doubles
    bool useGravityErrorTerms = false;
    bool useOdometryErrorTerms = false;
    const double* measuredGravity = nullptr;   // n x 3
    const int32_t* gravityPlausible = nullptr; // n
    const double* relativeTransl = nullptr;    // n x 3
    const double* relativeOrientMat = nullptr; // n x 9 row-major
    double balancingFactorGrav = 1.0;
    double balancingFactorOdom = 1000.0;
    double* relOrientations = nullptr;  // keyframePoses.relativePoses (3 x n, in/out)
    double* relTranslations = nullptr;
    double* globOrientations = nullptr;
    double* globTranslations = nullptr;
    float* globalPointsXYZW = nullptr;  // out N x 4
    float* globalNormalsXYZW = nullptr; // out N x 4
};

struct OptimReport {
    int iterations = 0;
    int stop_reason = 0;
    int num_gaussians = 0;
    double error0 = 0;
};

class DmsaOptimizerB200 {
    dmsa_b200_ctx* ctx_ = nullptr;

    void check(int rc, const char* what) {
        if (rc != DMSA_B200_OK) throw std::runtime_error(std::string("dmsa_b200: ") + what + ": " + dmsa_b200_last_error(ctx_));
    }
    static const char* message(int stop) {  // the reference's std::cout messages (DmsaOptimizer.h:91,119,132,141)
        switch (stop) {
            case DMSA_B200_STOP_FEW_GAUSSIANS: return "Number of gaussians is smaller than threshold, dmsa optimization is aborted";
            case DMSA_B200_STOP_NAN: return "Stop optimization because of NaN";
            case DMSA_B200_STOP_NO_IMPROVEMENT: return "Stop optimization because of no improvements";
            case DMSA_B200_STOP_EPSILON: return "Optimization step is smaller than epsilon";
            default: return "";
        }
    }

public:
    explicit DmsaOptimizerB200(int device = 0, void* cuda_stream = nullptr) {
        int rc = dmsa_b200_create(&ctx_, device, cuda_stream);
        if (rc != DMSA_B200_OK) throw std::runtime_error(rc == DMSA_B200_ERR_NO_DEVICE ? "dmsa_b200: no CUDA device (there is no CPU fallback)" : "dmsa_b200_create failed");
    }
    ~DmsaOptimizerB200() { dmsa_b200_destroy(ctx_); }
    DmsaOptimizerB200(const DmsaOptimizerB200&) = delete;
    DmsaOptimizerB200& operator=(const DmsaOptimizerB200&) = delete;

    bool verbose = true;  // print the reference's stop messages

    // optimizeSet on the sliding-window model (DmsaSlam.h:166)
    OptimReport optimizeSet(TrajectoryView& t, DmsaOptimSettings settings = DmsaOptimSettings()) {
        if (t.horizon > 0)
            check(dmsa_b200_traj_init_window(ctx_, t.t_min, t.horizon, t.numControlPoses, t.useImuErrorTerms, t.dt_res), "traj_init_window");
        else
            check(dmsa_b200_traj_init(ctx_, t.t_min, t.t_max, t.numControlPoses, t.useImuErrorTerms, t.dt_res), "traj_init");
        std::vector<const dmsa_b200_point_stamp_id*> ptr;
        std::vector<int64_t> sz;
        std::vector<float> gs;
        for (auto& s : t.scans) {
            ptr.push_back(s.points);
            sz.push_back(s.size);
            gs.push_back(s.gridSize);
        }
        check(dmsa_b200_traj_register_scans(ctx_, (int32_t)ptr.size(), ptr.data(), sz.data(), gs.data()), "register_scans");
        if (t.numStatic > 0) check(dmsa_b200_traj_add_static_points(ctx_, t.staticPoints, t.numStatic), "add_static_points");
        check(dmsa_b200_set_relative_poses(ctx_, t.relOrientations, t.relTranslations), "set_relative_poses");
        if (t.useImuErrorTerms) {
            if (!t.preintRot || !t.preintPos || !t.preintVel || !t.covInv)
                throw std::runtime_error("dmsa_b200: useImuErrorTerms is set but the preintegration factors are missing from the view");
            check(dmsa_b200_traj_set_imu_factors(ctx_, t.preintRot, t.preintPos, t.preintVel, t.covInv, t.balancingImu, t.gravity), "set_imu_factors");
        }
        dmsa_b200_settings c = to_c(settings);
        dmsa_b200_report rep;
        check(dmsa_b200_optimize(ctx_, &c, &rep), "optimize");
        check(dmsa_b200_get_poses(ctx_, t.relOrientations, t.relTranslations, t.globOrientations, t.globTranslations), "get_poses");
        if (t.globalPointsXYZW) check(dmsa_b200_get_global_points(ctx_, t.globalPointsXYZW, nullptr), "get_global_points");
        if (t.denseOrientations || t.denseTranslations) check(dmsa_b200_traj_get_dense_poses(ctx_, t.denseOrientations, t.denseTranslations), "get_dense_poses");
        if (t.denseTforms) check(dmsa_b200_traj_get_dense_tforms(ctx_, t.denseTforms), "get_dense_tforms");
        if (verbose && rep.stop_reason != DMSA_B200_STOP_MAX_ITER) std::printf("%s after iteration %d . . . \n", message(rep.stop_reason), rep.iterations - 1);
        return OptimReport{rep.iterations, rep.stop_reason, rep.num_gaussians, rep.error0};
    }

    // optimizeSet on the keyframe submap (DmsaSlam.h:228)
    OptimReport optimizeSet(SubmapView& m, DmsaOptimSettings settings = DmsaOptimSettings()) {
        const int n = (int)m.keyframes.size();
        check(dmsa_b200_kf_init(ctx_, n), "kf_init");
        for (int k = 0; k < n; ++k)
            check(dmsa_b200_kf_set_keyframe(ctx_, k, m.keyframes[k].points, m.keyframes[k].ringIds, m.keyframes[k].size, m.keyframes[k].gridSize), "kf_set_keyframe");
        check(dmsa_b200_kf_commit(ctx_), "kf_commit");
        check(dmsa_b200_set_relative_poses(ctx_, m.relOrientations, m.relTranslations), "set_relative_poses");
        if (m.useGravityErrorTerms) {  // MapManagement.h:162-190: the rows are [gravity | odometry]
            if (!m.measuredGravity || !m.gravityPlausible) throw std::runtime_error("dmsa_b200: useGravityErrorTerms is set but the gravity data are missing from the view");
            check(dmsa_b200_kf_set_gravity_terms(ctx_, m.measuredGravity, m.gravityPlausible, m.balancingFactorGrav), "kf_set_gravity_terms");
        }
        if (m.useOdometryErrorTerms) {
            if (!m.relativeTransl || !m.relativeOrientMat) throw std::runtime_error("dmsa_b200: useOdometryErrorTerms is set but the odometry data are missing from the view");
            check(dmsa_b200_kf_set_odometry_terms(ctx_, m.relativeTransl, m.relativeOrientMat, m.balancingFactorOdom), "kf_set_odometry_terms");
        }
        dmsa_b200_settings c = to_c(settings);
        dmsa_b200_report rep;
        check(dmsa_b200_optimize(ctx_, &c, &rep), "optimize");
        check(dmsa_b200_get_poses(ctx_, m.relOrientations, m.relTranslations, m.globOrientations, m.globTranslations), "get_poses");
        if (m.globalPointsXYZW) check(dmsa_b200_get_global_points(ctx_, m.globalPointsXYZW, m.globalNormalsXYZW), "get_global_points");
        if (verbose && rep.stop_reason != DMSA_B200_STOP_MAX_ITER) std::printf("%s after iteration %d . . . \n", message(rep.stop_reason), rep.iterations - 1);
        return OptimReport{rep.iterations, rep.stop_reason, rep.num_gaussians, rep.error0};
    }

    // ---- the nearest-neighbour step right before the sliding-window pass ----
    // DmsaSlam.h:304-339 for one keyframe cloud in the world frame, against the window cloud staged by the last optimizeSet /
    // update_global_points of this optimizer: selected[j] in {0, 1}; returns currOverlap.
    int64_t selectStaticPoints(const dmsa_b200_point_normal* cloud, int64_t n, const float pos[3], float maxDist, uint8_t* selected) {
        int64_t cnt = 0;
        check(dmsa_b200_select_static_points(ctx_, cloud, n, pos, maxDist, selected, &cnt), "select_static_points");
        return cnt;
    }
    // getOverlap(pc1, window cloud, maxDistOverlap), DmsaSlam.h:377-414
    float getOverlap(const float* pc1_xyzw, int64_t n1, float maxDistOverlap) {
        float ov = 0.0f;
        check(dmsa_b200_overlap(ctx_, pc1_xyzw, n1, maxDistOverlap, &ov), "overlap");
        return ov;
    }

    // ---- switches (every alternative is bit-identical) ----
    void setLmSolverOnDevice(bool on) { check(dmsa_b200_set_lm_solver(ctx_, on ? 0 : 1), "set_lm_solver"); }
    void setPairPackedKernels(bool on) { check(dmsa_b200_set_pair_mode(ctx_, on ? 1 : 0), "set_pair_mode"); }
    void setReferenceOrderMean(bool on) { check(dmsa_b200_set_mean_mode(ctx_, on ? 1 : 0), "set_mean_mode"); }
};

}  // namespace dmsa_b200

#ifdef DMSA_B200_WITH_REFERENCE_TYPES
// ---- binding against the reference's real types (compiled inside the reference's catkin workspace) ----------------
// #include "DMSA/DmsaOptimizer.h" (DmsaOptimSettings), "DMSA/ContinuousTrajectory.h" and "DMSA/MapManagement.h" before this
// header.  Only members the reference declares public are touched, by the names the reference gives them; Eigen objects are
// read through data() / operator() only.  tests/cpp/reference_mock.h carries mock classes with exactly these member names so
// that this branch is compiled and run (IMU, gravity and odometry factors included) without Eigen / PCL in the image.
static_assert(sizeof(PointStampId) == sizeof(dmsa_b200_point_stamp_id), "PointStampId layout (PointStampId.h:33-45)");
static_assert(sizeof(pcl::PointNormal) == sizeof(dmsa_b200_point_normal), "pcl::PointNormal layout");

template <typename PointT>
class DmsaOptimizerB200T {
    dmsa_b200::DmsaOptimizerB200 impl;

    // MapManagement's factor covariances are constants set in its constructor (MapManagement.h:66-70); the kernels carry the
    // same constants.  A caller that changed them gets an error instead of silently different residuals.
    template <class M3>
    static void requireScaledIdentity(const M3& C, double diag, const char* name) {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const double want = r == c ? diag : 0.0, got = C(r, c);
                if (!(std::fabs(got - want) <= 1e-9 * diag))
                    throw std::runtime_error(std::string("DmsaOptimizerB200: ") + name + " differs from the reference's constructor value (unsupported)");
            }
    }

public:
    explicit DmsaOptimizerB200T(int device = 0, void* cuda_stream = nullptr) : impl(device, cuda_stream) {}
    dmsa_b200::DmsaOptimizerB200& backend() { return impl; }

    void optimizeSet(OptimizablePointSet<PointT>& set, DmsaOptimSettings s = DmsaOptimSettings()) {
        dmsa_b200::DmsaOptimSettings c;  // field by field (DmsaOptimizer.h:25-39)
        c.num_iter = s.num_iter;
        c.epsilon = s.epsilon;
        c.use_analytic_jacobi = s.use_analytic_jacobi;
        c.step_length_optim = s.step_length_optim;
        c.max_step = s.max_step;
        c.gauss_split = s.gauss_split;
        c.grid_size_1_factor = s.grid_size_1_factor;
        c.grid_size_2_factor = s.grid_size_2_factor;
        c.min_num_points_per_set = s.min_num_points_per_set;
        c.min_num_gaussians = s.min_num_gaussians;
        c.lambda_diag = s.lambda_diag;
        c.use_centralization = s.use_centralization;
        if (auto* traj = dynamic_cast<ContinuousTrajectory*>(&set)) {
            dmsa_b200::TrajectoryView v;
            const int n = traj->controlPoses.numPoses;
            v.t_min = traj->t0;
            v.horizon = traj->horizon;  // what initTraj computed (:309), passed through unchanged
            v.t_max = traj->t0 + traj->horizon - traj->dt_res;
            v.dt_res = traj->dt_res;
            v.numControlPoses = n;
            v.useImuErrorTerms = traj->useImuErrorTerms;
            for (int k = 0; k < traj->regPcBuffer->getNumElements(); ++k) {
                PointCloudPlus& pc = traj->regPcBuffer->at(k);
                v.scans.push_back({reinterpret_cast<const dmsa_b200_point_stamp_id*>(pc.points.data()), (int64_t)pc.size(), pc.gridSize});
            }
            const int64_t nScan = traj->regPcBuffer->getNumPoints();
            const int64_t nAll = (int64_t)traj->globalPoints.points.size();
            v.staticPoints = reinterpret_cast<const dmsa_b200_point_stamp_id*>(traj->globalPoints.points.data() + nScan);
            v.numStatic = nAll - nScan;  // addStaticPoints appends them behind the scan points (:158-172)
            v.relOrientations = traj->controlPoses.relativePoses.Orientations.data();
            v.relTranslations = traj->controlPoses.relativePoses.Translations.data();
            v.globOrientations = traj->controlPoses.globalPoses.Orientations.data();
            v.globTranslations = traj->controlPoses.globalPoses.Translations.data();
            std::vector<float> xyzw(4 * (size_t)nAll);
            v.globalPointsXYZW = xyzw.data();
            // IMU factor constants (:36-43, 520-553): Eigen matrices are column-major, the C-ABI takes row-major
            std::vector<double> pr, pp, pv, ci;
            double grav[3] = {traj->gravity(0), traj->gravity(1), traj->gravity(2)};
            v.gravity = grav;
            v.balancingImu = traj->balancingImu;
            if (traj->useImuErrorTerms) {
                if ((int)traj->preintImuRots.size() < n || (int)traj->preintRelPositions.size() < n || (int)traj->preintRelVelocity.size() < n ||
                    (int)traj->CovPVRot_inv.size() < n)
                    throw std::runtime_error("DmsaOptimizerB200: useImuErrorTerms is set but the preintegration factors are not filled (updatePreintFactors)");
                pr.assign((size_t)n * 9, 0.0);
                pp.assign((size_t)n * 3, 0.0);
                pv.assign((size_t)n * 3, 0.0);
                ci.assign((size_t)n * 81, 0.0);
                for (int k = 1; k < n; ++k) {  // index 0 is never read (:617)
                    for (int r = 0; r < 3; ++r) {
                        for (int cc = 0; cc < 3; ++cc) pr[(size_t)k * 9 + 3 * r + cc] = traj->preintImuRots[k](r, cc);
                        pp[(size_t)k * 3 + r] = traj->preintRelPositions[k](r);
                        pv[(size_t)k * 3 + r] = traj->preintRelVelocity[k](r);
                    }
                    for (int r = 0; r < 9; ++r)
                        for (int cc = 0; cc < 9; ++cc) ci[(size_t)k * 81 + 9 * r + cc] = traj->CovPVRot_inv[k](r, cc);
                }
                v.preintRot = pr.data();
                v.preintPos = pp.data();
                v.preintVel = pv.data();
                v.covInv = ci.data();
            }
            // the dense poses / transforms of the final parameters are part of the set's state (getSubmapGravityEstimate reads
            // denseGlobalPoses after optimizeSet, :593-601)
            const int nt = traj->n_total;
            std::vector<float> dense(12 * (size_t)nt);
            v.denseOrientations = traj->denseGlobalPoses.Orientations.data();
            v.denseTranslations = traj->denseGlobalPoses.Translations.data();
            v.denseTforms = dense.data();
            impl.optimizeSet(v, c);
            for (int64_t i = 0; i < nAll; ++i) std::memcpy(traj->globalPoints.points[i].data, &xyzw[4 * (size_t)i], 16);
            for (int k = 0; k < nt; ++k) {  // Matrix4f, column-major; the last row stays (0, 0, 0, 1) (:316)
                float* M = traj->denseTformsLocal2Global[k].data();
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 4; ++cc) M[4 * cc + r] = dense[12 * (size_t)k + 4 * r + cc];
            }
        } else if (auto* map = dynamic_cast<MapManagement*>(&set)) {
            dmsa_b200::SubmapView v;
            const int n = map->keyframeDataBuffer.getNumElements();
            if (n != map->keyframePoses.numPoses)  // getSubmap builds MapManagement(nFrames) and fills all of them (MapManagement.h:254-276)
                throw std::runtime_error("DmsaOptimizerB200: keyframePoses.numPoses != keyframeDataBuffer.getNumElements() (pass a submap from getSubmap)");
            std::vector<double> mg((size_t)n * 3), rt((size_t)n * 3), rm((size_t)n * 9);
            std::vector<int32_t> pl(n);
            for (int k = 0; k < n; ++k) {
                auto& kf = map->keyframeDataBuffer.at(k);
                v.keyframes.push_back({reinterpret_cast<const dmsa_b200_point_normal*>(kf.pointCloudLocal->points.data()), kf.ringIds.data(),
                                       (int64_t)kf.pointCloudLocal->size(), kf.gridSize});
                pl[k] = kf.gravityPlausible ? 1 : 0;
                for (int r = 0; r < 3; ++r) {
                    mg[(size_t)k * 3 + r] = kf.measuredGravity(r);
                    rt[(size_t)k * 3 + r] = kf.relativeTransl(r);
                    for (int cc = 0; cc < 3; ++cc) rm[(size_t)k * 9 + 3 * r + cc] = kf.relativeOrientMat(r, cc);
                }
            }
            v.useGravityErrorTerms = map->useGravityErrorTerms;
            v.useOdometryErrorTerms = map->useOdometryErrorTerms;
            v.measuredGravity = mg.data();
            v.gravityPlausible = pl.data();
            v.relativeTransl = rt.data();
            v.relativeOrientMat = rm.data();
            v.balancingFactorGrav = map->balancingFactorGrav;
            v.balancingFactorOdom = map->balancingFactorOdom;
            if (map->useGravityErrorTerms) {
                requireScaledIdentity(map->Cov_grav_inv, 1.0 / (0.3 * 0.3), "Cov_grav_inv");
                if (!(map->gravity(0) == 0.0 && map->gravity(1) == 0.0 && map->gravity(2) == -9.805))
                    throw std::runtime_error("DmsaOptimizerB200: MapManagement::gravity differs from the reference's constructor value (unsupported)");
            }
            if (map->useOdometryErrorTerms) {
                requireScaledIdentity(map->odometryTranslCovInv, 1.0 / (0.01 * 0.01), "odometryTranslCovInv");
                requireScaledIdentity(map->odometryOrientCovInv, 1.0 / (0.01 * 0.01), "odometryOrientCovInv");
            }
            v.relOrientations = map->keyframePoses.relativePoses.Orientations.data();
            v.relTranslations = map->keyframePoses.relativePoses.Translations.data();
            v.globOrientations = map->keyframePoses.globalPoses.Orientations.data();
            v.globTranslations = map->keyframePoses.globalPoses.Translations.data();
            const size_t nAll = map->globalPoints.points.size();
            std::vector<float> xyzw(4 * nAll), nrm(4 * nAll);
            v.globalPointsXYZW = xyzw.data();
            v.globalNormalsXYZW = nrm.data();
            impl.optimizeSet(v, c);
            for (size_t i = 0; i < nAll; ++i) {
                std::memcpy(map->globalPoints.points[i].data, &xyzw[4 * i], 16);
                std::memcpy(map->globalPoints.points[i].data_n, &nrm[4 * i], 16);
            }
        } else {
            throw std::runtime_error("DmsaOptimizerB200: unsupported OptimizablePointSet subclass (no CPU fallback)");
        }
    }
};
#endif
