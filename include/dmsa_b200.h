/* ============================================================================
 * dmsa_b200.h — C-ABI of libdmsa_b200.so: the B200-native (sm_100a) DMSA inner loop.
 *
 * Drop-in boundary for ONE hot path of davidskdds/DMSA_LiDAR_SLAM (@2d20a58):
 *     DmsaOptimizer<PointT>::optimizeSet(OptimizablePointSet<PointT>&, DmsaOptimSettings)
 *     include/DMSA/DmsaOptimizer.h:54-150 and everything it calls per cost evaluation.
 * The reference has no FFI; its boundary is the C++ virtual interface
 * include/DMSA/OptimizablePointSet.h:18-56.  Because the optimizer calls back into the two
 * concrete point-set models on every cost evaluation (ContinuousTrajectory, MapManagement),
 * the GPU path absorbs the hot members of both models; the entry
 * points below are what a C++ adapter with the reference's own signature binds to
 * (see INTEGRATION.md and dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an int
 * status (0 = DMSA_B200_OK); no exceptions cross the ABI; host buffers are caller-owned,
 * device buffers are context-owned; one context per optimizer instance; a context is not
 * thread-safe (the reference optimizer is not re-entrant either: DmsaOptimizer.h:45-48).
 * Pose arrays are 3 x n column-major doubles exactly like Eigen::Matrix3Xd (Poses.h:19-20).
 * ========================================================================== */
#ifndef DMSA_B200_H
#define DMSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmsa_b200_ctx dmsa_b200_ctx;

enum {
    DMSA_B200_OK = 0,
    DMSA_B200_ERR_CUDA = 1,          /* a CUDA runtime call or kernel failed (see dmsa_b200_last_error) */
    DMSA_B200_ERR_ARG = 2,           /* invalid argument / call order */
    DMSA_B200_ERR_NO_DEVICE = 3,     /* no CUDA device: the product has NO CPU fallback */
    DMSA_B200_ERR_UNSUPPORTED = 4    /* input outside the supported envelope (e.g. homogeneous w != 1) */
};

/* Stop reasons of the optimisation loop (the reference prints a message and breaks). */
enum {
    DMSA_B200_STOP_MAX_ITER = 0,     /* loop ran num_iter iterations                       DmsaOptimizer.h:69   */
    DMSA_B200_STOP_FEW_GAUSSIANS = 1,/* numPointSets < min_num_gaussians                   DmsaOptimizer.h:89-93 */
    DMSA_B200_STOP_NAN = 2,          /* NaN in the step; parameters restored               DmsaOptimizer.h:116-122 */
    DMSA_B200_STOP_NO_IMPROVEMENT = 3,/* line search found nothing; set left at p+0.9*step DmsaOptimizer.h:130-134 */
    DMSA_B200_STOP_EPSILON = 4       /* ||step|| < epsilon                                  DmsaOptimizer.h:139-143 */
};

/* Field-for-field mirror of struct DmsaOptimSettings, DmsaOptimizer.h:25-39 (bool -> int32). */
typedef struct dmsa_b200_settings {
    int32_t num_iter;               /* = 15      */
    double epsilon;                 /* = 1e-5    */
    int32_t use_analytic_jacobi;    /* = false; never read by the reference (DmsaOptimizer.h:29) nor here */
    double step_length_optim;       /* = 0.05    */
    double max_step;                /* = 0.01    */
    int32_t gauss_split;            /* = false   */
    float grid_size_1_factor;       /* = 2.0     */
    float grid_size_2_factor;       /* = 5.0     */
    int32_t min_num_points_per_set; /* = 6       */
    int32_t min_num_gaussians;      /* = 30      */
    float lambda_diag;              /* = 0.00001 */
    int32_t use_centralization;     /* = true    */
} dmsa_b200_settings;

/* PointStampId, include/DMSA/PointStampId.h:33-45 (32 bytes, 16-aligned). */
typedef struct dmsa_b200_point_stamp_id {
    float x, y, z, w;
    double stamp;
    int32_t id;
    int32_t isStatic;
} dmsa_b200_point_stamp_id;

/* pcl::PointNormal (48 bytes): data[4], data_n[4], curvature, 3 floats padding. */
typedef struct dmsa_b200_point_normal {
    float x, y, z, w;
    float nx, ny, nz, nw;
    float curvature;
    float pad[3];
} dmsa_b200_point_normal;

/* Result of dmsa_b200_optimize / per-iteration trace. */
typedef struct dmsa_b200_report {
    int32_t iterations;      /* loop bodies executed (incl. the one that stopped) */
    int32_t stop_reason;     /* DMSA_B200_STOP_* */
    int32_t num_gaussians;   /* G of the last iteration */
    int64_t num_memberships; /* M of the last iteration */
    int32_t num_extra;       /* E additional residual rows */
    int32_t best_step;       /* line-search winner k (0..9) of the last iteration */
    double error0;           /* e0^T e0 of the last iteration */
    double step_norm;        /* ||clamped step||_2 of the last iteration */
} dmsa_b200_report;

/* ---- life cycle ------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t to launch on (e.g. torch's current stream) or NULL for an own stream. */
int dmsa_b200_create(dmsa_b200_ctx** out, int device, void* cuda_stream);
void dmsa_b200_destroy(dmsa_b200_ctx* ctx);
const char* dmsa_b200_last_error(const dmsa_b200_ctx* ctx);
int dmsa_b200_version(void);
/* sets with at most this many members are evaluated by the fused cost kernel (the rest by the chunked kernels) */
int32_t dmsa_b200_fuse_threshold(void);
/* number of kernels this library has launched on the context's stream since creation */
int64_t dmsa_b200_launch_count(const dmsa_b200_ctx* ctx);
int dmsa_b200_synchronize(dmsa_b200_ctx* ctx);

/* ---- sliding-window model: hot members of ContinuousTrajectory ---------------------------- */
/* initTraj(t_min, t_max, numControlPoses, useImu, dt_res)        ContinuousTrajectory.h:301-346 */
int dmsa_b200_traj_init(dmsa_b200_ctx* ctx, double t_min, double t_max, int32_t n_poses, int32_t use_imu, double dt_res);
/* the same, given the members initTraj leaves behind (t0, horizon = t_max - t_min + dt_res): what a binding passes that
 * sees an already initialised ContinuousTrajectory (ContinuousTrajectory.h:46-49), so `horizon` is not re-derived */
int dmsa_b200_traj_init_window(dmsa_b200_ctx* ctx, double t0, double horizon, int32_t n_poses, int32_t use_imu, double dt_res);
/* registerPcBuffer: scans in chronological ring-buffer order; computes tformIdPerPoint on device,
 * minGridSize = min(grid_sizes)                                  ContinuousTrajectory.h:228-261 */
int dmsa_b200_traj_register_scans(dmsa_b200_ctx* ctx, int32_t n_scans, const dmsa_b200_point_stamp_id* const* scans,
                                  const int64_t* sizes, const float* grid_sizes);
/* addStaticPoints / removeStaticPoints                            ContinuousTrajectory.h:158-187 */
int dmsa_b200_traj_add_static_points(dmsa_b200_ctx* ctx, const dmsa_b200_point_stamp_id* pts, int64_t n);
int dmsa_b200_traj_remove_static_points(dmsa_b200_ctx* ctx);
/* window timing as computed by initTraj (any pointer may be NULL) */
int dmsa_b200_traj_get_timing(dmsa_b200_ctx* ctx, int32_t* n_total, double* horizon, double* ctrl_stamps /*n_poses*/,
                              double* traj_time /*n_total*/, int32_t* param_indices /*n_poses*/);
/* per-point dense-pose index (tformIdPerPoint), n_scan_points entries */
int dmsa_b200_traj_get_tform_ids(dmsa_b200_ctx* ctx, int32_t* out);
/* IMU factor constants consumed by updateImuError                  ContinuousTrajectory.h:520-553, 603-663
 * preint_rot: n_poses x 9 row-major, preint_pos/vel: n_poses x 3, cov_inv: n_poses x 81 row-major (index 0 unused) */
int dmsa_b200_traj_set_imu_factors(dmsa_b200_ctx* ctx, const double* preint_rot, const double* preint_pos, const double* preint_vel,
                                   const double* cov_inv, double balancing_imu, const double* gravity3);

/* ---- keyframe-submap model: hot members of MapManagement ---------------------------------- */
int dmsa_b200_kf_init(dmsa_b200_ctx* ctx, int32_t n_keyframes);
/* keyframe k: local PointNormal cloud + ring ids + gridSize       KeyframeData.h:17-33 */
int dmsa_b200_kf_set_keyframe(dmsa_b200_ctx* ctx, int32_t k, const dmsa_b200_point_normal* pts, const int32_t* ring_ids, int64_t n,
                              float grid_size);
int dmsa_b200_kf_commit(dmsa_b200_ctx* ctx);
/* gravity / odometry factors                                      MapManagement.h:210-252 */
int dmsa_b200_kf_set_gravity_terms(dmsa_b200_ctx* ctx, const double* measured_gravity /*n x 3*/, const int32_t* plausible, double balance);
int dmsa_b200_kf_set_odometry_terms(dmsa_b200_ctx* ctx, const double* rel_transl /*n x 3*/, const double* rel_orient_mat /*n x 9 row-major*/,
                                    double balance);

/* ---- poses / parameters (both models) ------------------------------------------------------ */
/* relativePoses.Orientations / .Translations, 3 x n_poses column-major                  Poses.h:19-20 */
int dmsa_b200_set_relative_poses(dmsa_b200_ctx* ctx, const double* rel_orient, const double* rel_transl);
int dmsa_b200_get_poses(dmsa_b200_ctx* ctx, double* rel_orient, double* rel_transl, double* glob_orient, double* glob_transl);
int32_t dmsa_b200_num_params(const dmsa_b200_ctx* ctx);                 /* P = 6 (n_poses - 1)            Poses.h:59-62 */
int dmsa_b200_get_pose_parameters(dmsa_b200_ctx* ctx, double* params);  /* getParamsAsVector              Poses.h:64-70 */
int dmsa_b200_set_pose_parameters(dmsa_b200_ctx* ctx, const double* params); /* setParamsFromVector       Poses.h:72-76 */
int dmsa_b200_centralize(dmsa_b200_ctx* ctx);    /* ContinuousTrajectory.h:75-88  (no-op for keyframes: MapManagement.h:73-79) */
int dmsa_b200_decentralize(dmsa_b200_ctx* ctx);  /* ContinuousTrajectory.h:89-100 */

/* ---- the hot path, step by step ------------------------------------------------------------ */
/* updateGlobalPoints at the current parameters (device resident)   ContinuousTrajectory.h:129-156 | MapManagement.h:120-149 */
int dmsa_b200_update_global_points(dmsa_b200_ctx* ctx);
int64_t dmsa_b200_num_points(const dmsa_b200_ctx* ctx);
/* download globalPoints as N x 4 floats (xyzw); normals N x 4 (keyframe model only, may be NULL) */
int dmsa_b200_get_global_points(dmsa_b200_ctx* ctx, float* xyzw, float* normals);
/* dense local->global transforms of the current parameters: n_total x 12 floats (rows 0..2 of Matrix4f) */
int dmsa_b200_traj_get_dense_tforms(dmsa_b200_ctx* ctx, float* out);
/* validation surface: the float transform table of the batch evaluated last (after cost_jacobian: the forward-difference
 * batch [p, p + h e_0, ..]), (rows + 1) x V x 12 floats, row = dense sample / keyframe, the last row = identity (static
 * points); dims[2] = {rows + 1, V}.  out may be NULL (dimensions only). */
int dmsa_b200_get_batch_tables(dmsa_b200_ctx* ctx, float* out, int32_t* dims);
/* denseGlobalPoses of the current parameters (ContinuousTrajectory.h:29, 194-218): orientations (axis-angle) and
 * translations, 3 x n_total column-major doubles each (either may be NULL) */
int dmsa_b200_traj_get_dense_poses(dmsa_b200_ctx* ctx, double* orient, double* transl);

/* reset + createGaussianSets(res1) + createGaussianSets(res2) + updateRebalancingWeights on the current
 * global points                                   DmsaOptimizer.h:78-96, 275-350; Gaussians.h:130-201 */
int dmsa_b200_build_sets(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, int32_t* num_gaussians, int64_t* num_memberships);
/* debug getters (any pointer may be NULL): CSR offsets[G+1], members[M] (ascending point index per set), info[G*9] row-major,
 * weights[G], level[G] (0/1), key[G*3] (voxel key relative to the octree anchor), sub[G] (0 unsplit, 1/2 split halves).
 * Row order == the reference's leaf-iterator order. */
int dmsa_b200_get_sets(dmsa_b200_ctx* ctx, int64_t* offsets, int32_t* members, float* info, float* weights, int32_t* level, int32_t* key,
                       int32_t* sub);
/* per-point voxel keys of one resolution level (3 ints per point) + octree root origin/depth as PCL would hold them */
int dmsa_b200_get_voxel_keys(dmsa_b200_ctx* ctx, int32_t level, int32_t* keys /*N*3*/, int64_t* root_lo /*3*/, int32_t* depth);

/* updateErrorTerms for a batch of V parameter vectors (row-major V x P) -> e row-major V x (G+E)
 * one call == V cost evaluations of SURVEY §3.3                    DmsaOptimizer.h:234-273 */
int dmsa_b200_eval_cost(dmsa_b200_ctx* ctx, const double* params, int32_t n_vectors, double* e);
/* calcNumericJacobian + H = J^T J + g = J^T e0 at the current parameters.
 * H: P x P row-major (symmetric, WITHOUT lambda), g: P, err0 = e0^T e0; e0 (G+E) and J ((G+E) x P column-major like
 * Eigen::MatrixXd) are optional (NULL to skip)                      DmsaOptimizer.h:99-107, 199-232 */
int dmsa_b200_cost_jacobian(dmsa_b200_ctx* ctx, double* H, double* g, double* err0, double* e0, double* J);
/* one loop body (DmsaOptimizer.h:69-144) at the current state; *stop = DMSA_B200_STOP_* (0 = continue).
 * trace pointers optional: step[P] (clamped), ls_cost[9]. */
int dmsa_b200_iteration(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, int32_t* stop, dmsa_b200_report* report, double* step,
                        double* ls_cost);
/* the whole optimizeSet: centralize, loop, decentralize, final updateGlobalPoints   DmsaOptimizer.h:54-150 */
int dmsa_b200_optimize(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, dmsa_b200_report* report);

/* Per-set mean of the cost evaluation (DmsaOptimizer.h:249-254).  0 (default): order-free, exactly-rounded sum — what the
 * fast kernels compute; 1: the reference's sequential float accumulation in member order (validation mode: one block per
 * set, slow on sets with tens of thousands of members).  See DESIGN.md §3 "mean". */
int dmsa_b200_set_mean_mode(dmsa_b200_ctx* ctx, int32_t mode);

/* ---- per-kernel device timing (CUDA events on the context's stream; used by bench.py for the roofline) ---- */
int dmsa_b200_profile_enable(dmsa_b200_ctx* ctx, int32_t on); /* also resets the accumulators */
int32_t dmsa_b200_profile_num(void);
const char* dmsa_b200_profile_name(int32_t id);
int dmsa_b200_profile_read(dmsa_b200_ctx* ctx, int32_t id, double* total_ms, int64_t* count);

/* ---- multi-GPU row sharding: this rank owns the Gaussians g with g % world == rank ------------- */
int dmsa_b200_set_shard(dmsa_b200_ctx* ctx, int32_t rank, int32_t world);
/* device-resident variants for NCCL: pointers are DEVICE memory owned by the caller (e.g. torch tensors).
 * hg_dev: P*P + P + 1 doubles = [H | g | err0] partial sums over this rank's rows. */
int dmsa_b200_cost_jacobian_dev(dmsa_b200_ctx* ctx, double* hg_dev);
/* 9 partial line-search costs for step (host P doubles) into ls_dev[9] (device) */
int dmsa_b200_line_search_costs_dev(dmsa_b200_ctx* ctx, const double* step, double* ls_dev);

/* LM step of the keyframe-BUNDLE extension (BASELINE config 4): the all-reduced global system of the bundles is symmetric
 * positive definite and has no reference arithmetic to mirror, so step = -alpha (H + lambda I)^-1 g comes from a blocked
 * Cholesky factorisation on the device (one cooperative kernel), followed by the NaN guard and the infinity-norm clamp of
 * DmsaOptimizer.h:113-128.  hg = [H | g | err0]; tail[0] = err0, tail[1] = 0 ok / 1 NaN step / 2 not positive definite. */
int dmsa_b200_spd_solve_dev(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, const double* hg_dev, int32_t n_params, double* step_dev,
                            double* tail_dev);
int dmsa_b200_spd_solve(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, const double* hg, int32_t n_params, double* step, int32_t* flag);

/* Building blocks of one iteration over keyframe bundles (one context per bundle; every call is asynchronous on the
 * context's stream; all contexts of a rank must share ONE stream):
 *   bundle_jacobian   base transform, set build (deferred: no host wait), cost / forward-difference batch, J^T J, and the
 *                     scatter-add of [H_b | g_b | err0_b | #sets | #missed guesses] into the global buffer ghg_dev
 *                     (P_global^2 + P_global + 3 doubles) through idx_dev (P_local global parameter indices, device int32)
 *   all_reduce        NCCL sum over the ranks (communicator of dmsa_b200_comm_init)
 *   spd_solve_dev     the global LM step (above)
 *   bundle_line_search  9 trial costs of the bundle for the global step, added into gls_dev[9]
 *   bundle_verify     after the caller synchronised: set count, and whether a deferred build must be repeated (sync_build = 1) */
int dmsa_b200_bundle_jacobian(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, const int32_t* idx_dev, int32_t P_global, double* ghg_dev,
                              int32_t sync_build);
int dmsa_b200_bundle_line_search(dmsa_b200_ctx* ctx, const double* gstep_dev, const int32_t* idx_dev, double* gls_dev);
int dmsa_b200_bundle_verify(dmsa_b200_ctx* ctx, int32_t* num_gaussians, int32_t* redo);
int dmsa_b200_all_reduce(dmsa_b200_ctx* ctx, double* dev, int64_t count);

/* In-library exchange for dmsa_b200_iteration / dmsa_b200_optimize on a sharded context: one process per GPU, every rank
 * stages the SAME set and calls the same sequence; an iteration then all-reduces [H | g | err0] (P*P + P + 1 doubles) and the
 * 9 line-search costs with NCCL on the context's stream (intra-node NVLink / NVSwitch), and every rank takes the same step.
 * NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is torch's own copy).
 * Rank 0 creates the 128-byte id and hands it to the other ranks by any means (e.g. torch.distributed.broadcast over gloo). */
int dmsa_b200_comm_unique_id(void* id128);
int dmsa_b200_comm_init(dmsa_b200_ctx* ctx, const void* id128, int32_t rank, int32_t world); /* implies set_shard(rank, world) */
int dmsa_b200_comm_destroy(dmsa_b200_ctx* ctx);                                              /* back to set_shard(0, 1) */
int64_t dmsa_b200_collective_count(const dmsa_b200_ctx* ctx); /* NCCL all-reduces issued by this context */

/* Host-side LM step (DmsaOptimizer.h:107-128) on a host copy of the (all-reduced) [H | g | err0] buffer; no context needed.
 * explicit_inverse = 1: the reference's arithmetic, (-alpha * H.inverse()) * g with an LU inverse (what dmsa_b200_iteration uses);
 * explicit_inverse = 2: the same arithmetic with the inverse's columns spread over a few helper threads (bit-identical; what
 *                      dmsa_b200_iteration does while it waits for the read-back);
 * explicit_inverse = 0: Cholesky (LU fallback) solve of one right-hand side (used by the keyframe-bundle extension).
 * step: n_params doubles (clamped); *has_nan = 1 if the step contains NaN (the caller restores the parameters and stops). */
int dmsa_b200_lm_solve(const dmsa_b200_settings* settings, const double* hg, int32_t n_params, int32_t explicit_inverse, double* step,
                       int32_t* has_nan);

/* LM step of dmsa_b200_iteration / dmsa_b200_optimize (DmsaOptimizer.h:107-128):
 *   0 (default) device kernels for P <= 128 (kernels_solve.cuh: LU with partial pivoting + explicit inverse, the operation
 *     sequence of the host solver, bit-identical); larger systems use the host solver;
 *   1 host solver (host_solve.cpp), the reference's expression (-alpha * H.inverse()) * J^T e;
 *   2 device Cholesky (kernels_chol.cuh) for every P <= 1024: H + lambda I is symmetric positive definite, the step agrees
 *     with the LU-inverse step to the conditioning of the system (not bit for bit); a system the factorisation refuses
 *     falls back to the host solver.  Keeps loop bodies with P > 128 (keyframe submaps of more than 22 keyframes, BASELINE
 *     config 5) free of host round trips. */
int dmsa_b200_set_lm_solver(dmsa_b200_ctx* ctx, int32_t mode);
/* Cost kernels of the forward-difference batch (V = P + 1 vectors): 1 (default) = two vectors per thread with Blackwell's
 * packed FP32x2 instructions (FMUL2 / FADD2; every multiply->add edge keeps a scalar side, so nothing is fused) and the
 * shared-rotation fast path (the vectors that perturb a translation parameter reuse vector 0's rotated member coordinates:
 * their table rows carry the same rotation bits), 2 = pair-packed, every vector evaluated in full, 0 = one vector per
 * thread.  Same rounding sequence per value: bit-identical results in all three modes. */
int dmsa_b200_set_pair_mode(dmsa_b200_ctx* ctx, int32_t mode);
/* dmsa_b200_optimize with the device LM solver: 1 (default) = run-ahead loop (loop body i + 1 is enqueued before the host has
 * read body i's results: winner, parameter update and stop tests of DmsaOptimizer.h:113-143 run on the device, the host
 * consumes the read-back blocks one body late; a body that ran past a stop condition is discarded), 0 = one body at a time.
 * Bit-identical results. */
int dmsa_b200_set_run_ahead(dmsa_b200_ctx* ctx, int32_t on);
/* The device LM step on a HOST copy of [H | g | err0] (n_params <= 1024); validation twin of dmsa_b200_lm_solve(.., 1, ..). */
int dmsa_b200_lm_solve_device(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, const double* hg, int32_t n_params, double* step,
                              int32_t* has_nan);

/* ---- the nearest-neighbour step right before the sliding-window pass (DmsaSlam.h:264-414) ----
 * Both search the window cloud staged in this context (globalPoints as of the last update_global_points).
 * addStaticPoints, inner loop for ONE keyframe cloud in the world frame (DmsaSlam.h:304-339): selected[j] = 1 iff the nearest
 * window point is within max_dist (PCL KdTreeFLANN / flann::L2_Simple<float> squared distance <= pow(max_dist, 2), :293-315)
 * and the point is visible from pos (isVisible, :347-363).  *num_selected = currOverlap of that keyframe (:331). */
int dmsa_b200_select_static_points(dmsa_b200_ctx* ctx, const dmsa_b200_point_normal* cloud, int64_t n, const float* pos /*3*/, float max_dist,
                                   uint8_t* selected /*n*/, int64_t* num_selected);
/* getOverlap(pc1, pc2 = the window cloud, max_dist) (DmsaSlam.h:377-414): fraction of window points whose nearest pc1 point
 * (n1 x {x, y, z, w} floats on the host) is within max_dist; 0 when either cloud is empty. */
int dmsa_b200_overlap(dmsa_b200_ctx* ctx, const float* pc1_xyzw, int64_t n1, float max_dist, float* overlap);

/* ---- scan pre-processing and normal estimation (the producers of the optimizer's inputs) ----
 * The reference seeds every randomGridDownsampling call with srand(time(0)); here the caller passes the seed, and the draw
 * is glibc's rand() sequence for that seed (dmsa_b200_rand_sequence), one number per octree leaf in leaf order. */
typedef struct dmsa_b200_preprocess_config {
    int32_t max_num_points_per_scan; /* Config.h:24 */
    float min_dist_ds;               /* Config.h:25 minDistDS */
    float min_dist;                  /* Config.h:38 */
    float lidar_to_imu[16];          /* Config.h:58 lidarToImuTform, column-major (Eigen::Matrix4f::data()) */
} dmsa_b200_preprocess_config;
/* out[k] = the k-th rand() after srand(seed) (glibc; host only, no context) */
int dmsa_b200_rand_sequence(uint32_t seed, int64_t n, int32_t* out);
/* randomGridDownsampling(rawPc, filteredPc, gridSize) (helpers.h:67-182): points = n records of stride_bytes (a multiple of
 * 16) with x, y, z floats at offset 0 (PointStampId: 32, pcl::PointNormal: 48); indices_out[c] = index of the point copied to
 * filteredPc->points[c] (PCL octree leaves in depth-first order, member int(r * (size - 1)) of each); *n_out = leaf count. */
int dmsa_b200_grid_downsample(dmsa_b200_ctx* ctx, const void* points, int64_t n, int32_t stride_bytes, float grid_size, uint32_t seed,
                              int32_t* indices_out /*n*/, int64_t* n_out);
/* the same on the staged window's globalPoints (addNewKeyframeToMap, DmsaSlam.h:506) */
int dmsa_b200_downsample_global_points(dmsa_b200_ctx* ctx, float grid_size, uint32_t seed, int32_t* indices_out /*num_points*/, int64_t* n_out);
/* preProcess(rawPc, filteredPc) (DmsaSlam.h:570-634): adaptive grid (0.4 / 0.3 / 0.2 / 0.15 while fewer than max_num points),
 * range cut at max(rangesSorted[min(max_num, size - 1)], minDistDS) and > min_dist, transform to the IMU frame (the arithmetic
 * of pcl::transformPointCloud, PCL 1.10 SSE2 path), w = 1.  out: room for n records. */
int dmsa_b200_preprocess_scan(dmsa_b200_ctx* ctx, const dmsa_b200_point_stamp_id* raw, int64_t n, const dmsa_b200_preprocess_config* cfg, uint32_t seed,
                              dmsa_b200_point_stamp_id* out, int64_t* n_out, float* grid_size_out);
/* updateNormals(cloud, origin) (DmsaSlam.h:557-568): pcl::NormalEstimationOMP with setKSearch(6) on the cloud itself, normals
 * flipped towards the view point; normal_x/y/z and curvature are overwritten in place.  cell_size > 0: edge of the search
 * grid (the cloud's grid size).  nn_indices (optional, n x 6): neighbour indices in search-result order (ascending distance,
 * the point itself first), -1 where the cloud has fewer than 6 points. */
int dmsa_b200_estimate_normals(dmsa_b200_ctx* ctx, dmsa_b200_point_normal* cloud, int64_t n, const float* viewpoint /*3*/, float cell_size,
                               int32_t* nn_indices);

/* ---- the data formats either side of the path ------------------------------------------------
 * sensor_msgs/PointCloud2 -> PointStampId (src/dmsa_slam_ros.cpp:372-486): where the x / y / z / time / ring fields sit
 * inside one point record and how the reference interprets them per sensor type. */
enum { DMSA_B200_STAMP_NONE = 0,       /* "unknown": stampMsg + deltaT * k / n                                  */
       DMSA_B200_STAMP_F64_ABS = 1,    /* hesai, robosense, livoxXYZRTLT_s: double seconds                     */
       DMSA_B200_STAMP_U32_NS_REL = 2, /* ouster: stampMsg + 1e-9 * uint32 nanoseconds                         */
       DMSA_B200_STAMP_F32_REL = 3,    /* velodyne, sick: stampMsg + float seconds                             */
       DMSA_B200_STAMP_F64_NS_ABS = 4  /* livoxXYZRTLT_ns: 1e-9 * double nanoseconds                           */ };
enum { DMSA_B200_RING_NONE = 0 /* artificial ring k % 1000 */, DMSA_B200_RING_U16 = 1, DMSA_B200_RING_U8 = 2, DMSA_B200_RING_I8 = 3 };
typedef struct dmsa_b200_pc2_layout {
    int32_t point_step;
    int32_t x_offset, y_offset, z_offset;
    int32_t stamp_offset, stamp_type;
    int32_t ring_offset, ring_type;
} dmsa_b200_pc2_layout;
/* layout of one of the reference's sensor types ("hesai", "ouster", "robosense", "velodyne", "livoxXYZRTLT_s",
 * "livoxXYZRTLT_ns", "sick", "unknown") given msg->fields[i].offset for i < n_fields (host only, no context) */
int dmsa_b200_pc2_layout_for_sensor(const char* sensor, const int32_t* field_offsets, int32_t n_fields, int32_t point_step, dmsa_b200_pc2_layout* out);
/* the per-point loop of callbackPointCloud: data = msg->data (host), out = n_points records (host); stamp_msg =
 * msg->header.stamp.toSec(), delta_t = stampMsg - lastPcMsgStamp (only read for DMSA_B200_STAMP_NONE) */
int dmsa_b200_decode_pointcloud2(dmsa_b200_ctx* ctx, const uint8_t* data, int64_t n_points, const dmsa_b200_pc2_layout* layout, double stamp_msg,
                                 double delta_t, dmsa_b200_point_stamp_id* out);
/* ConsecutivePoses::relative2global (ConsecutivePoses.h:26-43) on 3 x n column-major arrays (host only, no context) */
int dmsa_b200_relative2global(int32_t n, const double* rel_orient, const double* rel_transl, double* glob_orient, double* glob_transl);
/* addPoseToFile (OutputManagement.h:80-96): one line of the TUM trajectory file; returns the length, -1 if buf is too small */
int dmsa_b200_format_tum_pose(double stamp, const double* pos /*3*/, const double* orient /*3, axis-angle*/, char* buf, int32_t buf_size);
/* pcl::io::savePCDFileASCII of a pcl::PointNormal cloud (src/dmsa_slam_ros.cpp:286-291); 0 on success, -1 on I/O failure */
int dmsa_b200_save_pcd_ascii(const char* filename, const dmsa_b200_point_normal* cloud, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* DMSA_B200_H */
