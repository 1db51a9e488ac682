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
        if (t.useImuErrorTerms)
            check(dmsa_b200_traj_set_imu_factors(ctx_, t.preintRot, t.preintPos, t.preintVel, t.covInv, t.balancingImu, nullptr), "set_imu_factors");
        dmsa_b200_settings c = to_c(settings);
        dmsa_b200_report rep;
        check(dmsa_b200_optimize(ctx_, &c, &rep), "optimize");
        check(dmsa_b200_get_poses(ctx_, t.relOrientations, t.relTranslations, t.globOrientations, t.globTranslations), "get_poses");
        if (t.globalPointsXYZW) check(dmsa_b200_get_global_points(ctx_, t.globalPointsXYZW, nullptr), "get_global_points");
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
// #include "DMSA/ContinuousTrajectory.h" / "DMSA/MapManagement.h" before this header.
static_assert(sizeof(PointStampId) == sizeof(dmsa_b200_point_stamp_id), "PointStampId layout (PointStampId.h:33-45)");
static_assert(sizeof(pcl::PointNormal) == sizeof(dmsa_b200_point_normal), "pcl::PointNormal layout");

template <typename PointT>
class DmsaOptimizerB200T {
    dmsa_b200::DmsaOptimizerB200 impl;

public:
    void optimizeSet(OptimizablePointSet<PointT>& set, DmsaOptimSettings s = DmsaOptimSettings()) {
        dmsa_b200::DmsaOptimSettings c;
        std::memcpy(&c, &s, sizeof(c));  // same fields in the same order
        if (auto* traj = dynamic_cast<ContinuousTrajectory*>(&set)) {
            dmsa_b200::TrajectoryView v;
            v.t_min = traj->t0;
            v.t_max = traj->t0 + traj->horizon - traj->dt_res;  // initTraj: horizon = t_max - t_min + dt_res (:309)
            v.dt_res = traj->dt_res;
            v.numControlPoses = traj->controlPoses.numPoses;
            v.useImuErrorTerms = traj->useImuErrorTerms;
            for (int k = 0; k < traj->regPcBuffer->getNumElements(); ++k) {
                PointCloudPlus& pc = traj->regPcBuffer->at(k);
                v.scans.push_back({reinterpret_cast<const dmsa_b200_point_stamp_id*>(pc.points.data()), (int64_t)pc.size(), pc.gridSize});
            }
            int64_t nScan = traj->regPcBuffer->getNumPoints();
            v.staticPoints = reinterpret_cast<const dmsa_b200_point_stamp_id*>(traj->globalPoints.points.data() + nScan);
            v.numStatic = (int64_t)traj->globalPoints.points.size() - nScan;
            v.relOrientations = traj->controlPoses.relativePoses.Orientations.data();
            v.relTranslations = traj->controlPoses.relativePoses.Translations.data();
            v.globOrientations = traj->controlPoses.globalPoses.Orientations.data();
            v.globTranslations = traj->controlPoses.globalPoses.Translations.data();
            std::vector<float> xyzw(4 * traj->globalPoints.points.size());
            v.globalPointsXYZW = xyzw.data();
            // IMU constants: flatten preintImuRots / preintRelPositions / preintRelVelocity / CovPVRot_inv row-major here
            impl.optimizeSet(v, c);
            for (size_t i = 0; i < traj->globalPoints.points.size(); ++i) std::memcpy(traj->globalPoints.points[i].data, &xyzw[4 * i], 16);
        } else if (auto* map = dynamic_cast<MapManagement*>(&set)) {
            dmsa_b200::SubmapView v;
            for (int k = 0; k < map->keyframeDataBuffer.getNumElements(); ++k) {
                auto& kf = map->keyframeDataBuffer.at(k);
                v.keyframes.push_back({reinterpret_cast<const dmsa_b200_point_normal*>(kf.pointCloudLocal->points.data()), kf.ringIds.data(),
                                       (int64_t)kf.pointCloudLocal->size(), kf.gridSize});
            }
            v.relOrientations = map->keyframePoses.relativePoses.Orientations.data();
            v.relTranslations = map->keyframePoses.relativePoses.Translations.data();
            v.globOrientations = map->keyframePoses.globalPoses.Orientations.data();
            v.globTranslations = map->keyframePoses.globalPoses.Translations.data();
            std::vector<float> xyzw(4 * map->globalPoints.points.size()), nrm(4 * map->globalPoints.points.size());
            v.globalPointsXYZW = xyzw.data();
            v.globalNormalsXYZW = nrm.data();
            impl.optimizeSet(v, c);
            for (size_t i = 0; i < map->globalPoints.points.size(); ++i) {
                std::memcpy(map->globalPoints.points[i].data, &xyzw[4 * i], 16);
                std::memcpy(map->globalPoints.points[i].data_n, &nrm[4 * i], 16);
            }
        } else {
            throw std::runtime_error("DmsaOptimizerB200: unsupported OptimizablePointSet subclass (no CPU fallback)");
        }
    }
};
#endif
