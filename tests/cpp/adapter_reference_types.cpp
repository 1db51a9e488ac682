// adapter_reference_types.cpp — compiles and runs the DMSA_B200_WITH_REFERENCE_TYPES branch of the C++ adapter
// (DmsaOptimizerB200T<PointT>::optimizeSet(OptimizablePointSet<PointT>&, DmsaOptimSettings), the drop-in for
// DmsaSlam.h:52-53, 166, 228) against mock classes with the reference's member names (reference_mock.h).
// Input: a dump written by tests/test_cpp_adapter.py (sliding window with IMU factors, or keyframe submap with gravity /
// odometry factors); output: the mutated set (poses, globalPoints, dense poses / transforms).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "reference_mock.h"
#define DMSA_B200_WITH_REFERENCE_TYPES
#include "../../dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h"

static FILE* fin;
template <class T>
static T rd() {
    T v;
    if (fread(&v, sizeof(T), 1, fin) != 1) {
        std::fprintf(stderr, "short read\n");
        std::exit(2);
    }
    return v;
}
static void rdn(void* p, size_t n) {
    if (n && fread(p, 1, n, fin) != n) {
        std::fprintf(stderr, "short read\n");
        std::exit(2);
    }
}
static DmsaOptimSettings readSettings() {
    DmsaOptimSettings s;
    s.num_iter = rd<int32_t>();
    s.step_length_optim = rd<double>();
    s.max_step = rd<double>();
    s.gauss_split = rd<int32_t>() != 0;
    s.min_num_points_per_set = rd<int32_t>();
    s.min_num_gaussians = rd<int32_t>();
    s.epsilon = rd<double>();
    return s;
}

static int runTrajectory(const char* outPath) {
    ContinuousTrajectory traj;
    const int n_scans = rd<int32_t>(), n = rd<int32_t>(), use_imu = rd<int32_t>();
    const int64_t n_static = rd<int64_t>();
    const double t_min = rd<double>(), t_max = rd<double>(), dt_res = rd<double>();
    DmsaOptimSettings s = readSettings();
    // initTraj (ContinuousTrajectory.h:301-346): the members it leaves behind
    traj.dt_res = dt_res;
    traj.useImuErrorTerms = use_imu != 0;
    traj.t0 = t_min;
    traj.horizon = t_max - t_min + dt_res;
    traj.n_total = (int)std::round(traj.horizon / dt_res) + 1;
    traj.denseTformsLocal2Global.resize(traj.n_total);
    for (auto& M : traj.denseTformsLocal2Global)
        for (int i = 0; i < 4; ++i) M(i, i) = 1.0f;
    traj.denseGlobalPoses.resize(traj.n_total);
    traj.numParams = n;
    traj.controlPoses = StampedConsecutivePoses(n);
    traj.preintImuRots.resize(n);
    traj.preintRelPositions.resize(n);
    traj.preintRelVelocity.resize(n);
    traj.CovPVRot_inv.resize(n);
    traj.gravity(0) = 0.0;
    traj.gravity(1) = 0.0;
    traj.gravity(2) = -9.805;
    // registerPcBuffer (:228-261) + addStaticPoints (:158-172)
    traj.regPcBuffer = std::make_shared<PointCloudBuffer>();
    for (int k = 0; k < n_scans; ++k) {
        PointCloudPlus pc;
        const int64_t m = rd<int64_t>();
        pc.gridSize = rd<float>();
        pc.points.resize(m);
        rdn(pc.points.data(), m * sizeof(PointStampId));
        traj.regPcBuffer->addElem(pc);
    }
    for (int k = 0; k < n_scans; ++k)
        for (auto& p : traj.regPcBuffer->at(k).points) traj.globalPoints.points.push_back(p);
    {
        std::vector<PointStampId> st(n_static);
        rdn(st.data(), n_static * sizeof(PointStampId));
        for (auto& p : st) traj.globalPoints.points.push_back(p);
    }
    rdn(traj.controlPoses.relativePoses.Orientations.data(), 3 * n * 8);
    rdn(traj.controlPoses.relativePoses.Translations.data(), 3 * n * 8);
    if (use_imu) {  // row-major in the dump -> the Eigen-like column-major members
        std::vector<double> pr(9 * n), pp(3 * n), pv(3 * n), ci(81 * (size_t)n);
        rdn(pr.data(), pr.size() * 8);
        rdn(pp.data(), pp.size() * 8);
        rdn(pv.data(), pv.size() * 8);
        rdn(ci.data(), ci.size() * 8);
        traj.balancingImu = rd<double>();
        for (int k = 0; k < n; ++k) {
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) traj.preintImuRots[k](r, c) = pr[9 * k + 3 * r + c];
                traj.preintRelPositions[k](r) = pp[3 * k + r];
                traj.preintRelVelocity[k](r) = pv[3 * k + r];
            }
            for (int r = 0; r < 9; ++r)
                for (int c = 0; c < 9; ++c) traj.CovPVRot_inv[k](r, c) = ci[81 * (size_t)k + 9 * r + c];
        }
    }
    std::fclose(fin);
    DmsaOptimizerB200T<PointStampId> slidingWindowOptimizer;  // DmsaSlam.h:52
    slidingWindowOptimizer.backend().verbose = false;
    slidingWindowOptimizer.optimizeSet(traj, s);              // DmsaSlam.h:166
    FILE* o = std::fopen(outPath, "wb");
    std::fwrite(traj.controlPoses.relativePoses.Orientations.data(), 8, 3 * n, o);
    std::fwrite(traj.controlPoses.relativePoses.Translations.data(), 8, 3 * n, o);
    std::fwrite(traj.controlPoses.globalPoses.Orientations.data(), 8, 3 * n, o);
    std::fwrite(traj.controlPoses.globalPoses.Translations.data(), 8, 3 * n, o);
    for (auto& p : traj.globalPoints.points) std::fwrite(p.data, 4, 4, o);
    std::fwrite(traj.denseGlobalPoses.Orientations.data(), 8, 3 * (size_t)traj.n_total, o);
    std::fwrite(traj.denseGlobalPoses.Translations.data(), 8, 3 * (size_t)traj.n_total, o);
    for (auto& M : traj.denseTformsLocal2Global) std::fwrite(M.data(), 4, 16, o);
    std::fclose(o);
    return 0;
}

static int runSubmap(const char* outPath) {
    const int n = rd<int32_t>(), use_grav = rd<int32_t>(), use_odom = rd<int32_t>();
    DmsaOptimSettings s = readSettings();
    MapManagement map(n);
    map.useGravityErrorTerms = use_grav != 0;
    map.useOdometryErrorTerms = use_odom != 0;
    map.balancingFactorGrav = rd<double>();
    map.balancingFactorOdom = rd<double>();
    size_t total = 0;
    for (int k = 0; k < n; ++k) {
        KeyframeData kf;
        const int64_t m = rd<int64_t>();
        kf.gridSize = rd<float>();
        kf.pointCloudLocal = std::make_shared<pcl::PointCloud<pcl::PointNormal>>();
        kf.pointCloudLocal->points.resize(m);
        rdn(kf.pointCloudLocal->points.data(), m * sizeof(pcl::PointNormal));
        kf.ringIds.resize((int)m);
        rdn(kf.ringIds.data(), m * 4);
        rdn(kf.measuredGravity.data(), 24);
        kf.gravityPlausible = rd<int32_t>() != 0;
        rdn(kf.relativeTransl.data(), 24);
        double R[9];
        rdn(R, 72);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) kf.relativeOrientMat(r, c) = R[3 * r + c];
        map.keyframeDataBuffer.addElem(kf);
        total += m;
    }
    map.globalPoints.resize(total);  // MapManagement.h:369
    rdn(map.keyframePoses.relativePoses.Orientations.data(), 3 * n * 8);
    rdn(map.keyframePoses.relativePoses.Translations.data(), 3 * n * 8);
    std::fclose(fin);
    DmsaOptimizerB200T<pcl::PointNormal> keyframeMapOptimizer;  // DmsaSlam.h:53
    keyframeMapOptimizer.backend().verbose = false;
    keyframeMapOptimizer.optimizeSet(map, s);                   // DmsaSlam.h:228
    FILE* o = std::fopen(outPath, "wb");
    std::fwrite(map.keyframePoses.relativePoses.Orientations.data(), 8, 3 * n, o);
    std::fwrite(map.keyframePoses.relativePoses.Translations.data(), 8, 3 * n, o);
    std::fwrite(map.keyframePoses.globalPoses.Orientations.data(), 8, 3 * n, o);
    std::fwrite(map.keyframePoses.globalPoses.Translations.data(), 8, 3 * n, o);
    for (auto& p : map.globalPoints.points) {
        std::fwrite(p.data, 4, 4, o);
        std::fwrite(p.data_n, 4, 4, o);
    }
    std::fclose(o);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    fin = std::fopen(argv[1], "rb");
    if (!fin) return 1;
    const int kind = rd<int32_t>();
    try {
        return kind == 0 ? runTrajectory(argv[2]) : runSubmap(argv[2]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
