// adapter_smoke.cpp — exercises the header-only C++ adapter (dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h) over the C-ABI.
// Reads a window dumped by tests/test_cpp_adapter.py, runs optimizeSet like DmsaSlam.h:166 would, writes the poses back.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h"

static void rd(FILE* f, void* p, size_t n) {
    if (fread(p, 1, n, f) != n) {
        std::fprintf(stderr, "short read\n");
        std::exit(2);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t n_scans, n_poses, num_iter;
    int64_t n_static;
    double t_min, t_max, dt_res;
    rd(f, &n_scans, 4);
    rd(f, &n_poses, 4);
    rd(f, &num_iter, 4);
    rd(f, &n_static, 8);
    rd(f, &t_min, 8);
    rd(f, &t_max, 8);
    rd(f, &dt_res, 8);
    std::vector<std::vector<dmsa_b200_point_stamp_id>> scans(n_scans);
    dmsa_b200::TrajectoryView v;
    for (int s = 0; s < n_scans; ++s) {
        int64_t n;
        float gs;
        rd(f, &n, 8);
        rd(f, &gs, 4);
        scans[s].resize(n);
        rd(f, scans[s].data(), n * sizeof(dmsa_b200_point_stamp_id));
        v.scans.push_back({scans[s].data(), n, gs});
    }
    std::vector<dmsa_b200_point_stamp_id> stat(n_static);
    rd(f, stat.data(), n_static * sizeof(dmsa_b200_point_stamp_id));
    std::vector<double> ro(3 * n_poses), rt(3 * n_poses), go(3 * n_poses), gt(3 * n_poses);
    rd(f, ro.data(), ro.size() * 8);
    rd(f, rt.data(), rt.size() * 8);
    std::fclose(f);
    v.t_min = t_min;
    v.t_max = t_max;
    v.dt_res = dt_res;
    v.numControlPoses = n_poses;
    v.staticPoints = stat.data();
    v.numStatic = n_static;
    v.relOrientations = ro.data();
    v.relTranslations = rt.data();
    v.globOrientations = go.data();
    v.globTranslations = gt.data();
    dmsa_b200::DmsaOptimSettings s;
    s.num_iter = num_iter;
    s.step_length_optim = 0.2;
    s.max_step = 0.3;
    s.min_num_points_per_set = 6;
    s.min_num_gaussians = 10;
    try {
        dmsa_b200::DmsaOptimizerB200 opt(0);
        dmsa_b200::OptimReport r = opt.optimizeSet(v, s);
        FILE* o = std::fopen(argv[2], "wb");
        std::fwrite(&r.iterations, 4, 1, o);
        std::fwrite(&r.stop_reason, 4, 1, o);
        std::fwrite(ro.data(), 8, ro.size(), o);
        std::fwrite(rt.data(), 8, rt.size(), o);
        std::fwrite(gt.data(), 8, gt.size(), o);
        std::fclose(o);
        std::printf("adapter_smoke: %d iterations, stop %d, G %d\n", r.iterations, r.stop_reason, r.num_gaussians);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return 0;
}
