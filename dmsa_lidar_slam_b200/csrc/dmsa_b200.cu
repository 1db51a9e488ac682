// dmsa_b200.cu — C-ABI implementation (include/dmsa_b200.h): context, HBM staging, the optimizeSet loop.
//
// Host control flow restates DmsaOptimizer.h:54-182 (optimizeSet, adaptiveStepSize); all per-point and
// per-set work runs in the sm_100a kernels of kernels_pose.cuh / kernels_sets.cuh / kernels_cost.cuh.
// There is NO CPU fallback: without a CUDA device dmsa_b200_create fails with DMSA_B200_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/dmsa_b200.h"
#include "kernels_chol.cuh"
#include "kernels_cost.cuh"
#include "kernels_solve.cuh"
#include "kernels_knn.cuh"
#include "kernels_pose.cuh"
#include "kernels_pre.cuh"
#include "kernels_sets.cuh"
#include "kernels_sort.cuh"
#include "nccl_dyn.h"
#include "se3_math.cuh"

#include <atomic>

// ---- programmatic dependent launch bookkeeping (pdl.cuh) ---------------------------------------------------------------
// A kernel may only carry the programmatic-stream-serialization attribute when the operation right before it ON ITS
// STREAM is a kernel of this library: griddepcontrol.wait orders a kernel behind the previous GRID only, and a kernel
// launched with the attribute behind a copy / memset / event wait was measured to start before that operation had
// finished (pageable uploads of the front-end calls, parameter uploads of the keyframe pass).  Every such stream
// operation of the library therefore bumps a process-wide epoch, a launch chains to its predecessor only when no such
// operation was issued since that predecessor, and only inside a PdlScope (a span of a C-ABI call in which nothing but
// this library enqueues work on the context's streams; the first launch of an outermost scope never chains).
static std::atomic<unsigned long long> g_opEpoch{1};
// Once the process owns an NCCL communicator no launch chains any more: NCCL enqueues its own kernels on the contexts'
// streams (with whatever launch attributes its version uses), and the multi-GPU runs that would have to prove the
// combination are the expensive ones.  The sharded paths therefore launch exactly as they did before PDL.
static std::atomic<bool> g_commSeen{false};
static inline void pdlBreak() { g_opEpoch.fetch_add(1, std::memory_order_relaxed); }
#define cudaMemcpyAsync(...) (pdlBreak(), cudaMemcpyAsync(__VA_ARGS__))
#define cudaMemsetAsync(...) (pdlBreak(), cudaMemsetAsync(__VA_ARGS__))
#define cudaMemcpy(...) (pdlBreak(), cudaMemcpy(__VA_ARGS__))
#define cudaMemcpyFromSymbol(...) (pdlBreak(), cudaMemcpyFromSymbol(__VA_ARGS__))
#define cudaStreamWaitEvent(...) (pdlBreak(), cudaStreamWaitEvent(__VA_ARGS__))
// (cudaEventRecord does not break a chain: the record completes with the kernel before it, and the kernel behind it waits for
// that same kernel in DMSA_PDL_ENTER() whether or not the driver lets it start early)
#define cudaLaunchCooperativeKernel(...) (pdlBreak(), cudaLaunchCooperativeKernel(__VA_ARGS__))

using namespace dmsa;

namespace {

constexpr int CHUNK = COST_CHUNK;     // members per work unit of the chunked cost kernels (big sets)
constexpr int FUSE_MAX = COST_CHUNK;  // sets up to this many members are evaluated by one block of the fused cost kernel

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostPoses {
    int n = 0;
    std::vector<double> relO, relT, globO, globT;  // 3 x n column-major
    void resize(int n_) {
        n = n_;
        relO.assign(3 * n, 0.0);
        relT.assign(3 * n, 0.0);
        globO.assign(3 * n, 0.0);
        globT.assign(3 * n, 0.0);
    }
    static Vec3 col(const std::vector<double>& a, int k) { return mk3(a[3 * k], a[3 * k + 1], a[3 * k + 2]); }
    static void setcol(std::vector<double>& a, int k, const Vec3& v) {
        a[3 * k] = v.x;
        a[3 * k + 1] = v.y;
        a[3 * k + 2] = v.z;
    }
    void relative2global() {  // ConsecutivePoses.h:26-43
        Mat3 R = identity3();
        Vec3 T = mk3(0, 0, 0);
        for (int k = 0; k < n; ++k) {
            T = add3(T, matvec3(R, col(relT, k)));
            setcol(globT, k, T);
            R = matmul3(R, so3_exp(col(relO, k)));
            setcol(globO, k, so3_log(R));
        }
    }
    void global2relative() {  // ConsecutivePoses.h:45-67
        setcol(relO, 0, col(globO, 0));
        setcol(relT, 0, col(globT, 0));
        for (int k = n - 1; k > 0; --k) {
            Mat3 R1 = so3_exp(col(globO, k - 1)), R2 = so3_exp(col(globO, k));
            setcol(relO, k, so3_log(matmul3(transpose3(R1), R2)));
            setcol(relT, k, matvec3(transpose3(R1), sub3(col(globT, k), col(globT, k - 1))));
        }
    }
    void getParams(std::vector<double>& p) const {  // Poses.h:64-70
        int m = 3 * (n - 1);
        p.resize(2 * (size_t)m);
        for (int i = 0; i < m; ++i) {
            p[i] = relO[3 + i];
            p[m + i] = relT[3 + i];
        }
    }
    void setParams(const double* p) {  // Poses.h:72-76
        int m = 3 * (n - 1);
        for (int i = 0; i < m; ++i) {
            relO[3 + i] = p[i];
            relT[3 + i] = p[m + i];
        }
    }
};

// dense P x P solvers of the LM step live in host_solve.cpp (plain C++, compiled by the host compiler with
// per-ISA clones: AVX-512 / AVX2 / baseline — element-wise IEEE arithmetic, identical results on every path)
}  // namespace
bool dmsa_host_lu_inverse(const std::vector<double>& A, int n, std::vector<double>& inv);
bool dmsa_host_lm_step(const double* hg, int n, double lambda, double alpha, double* step);
void dmsa_host_solver_arm();
void dmsa_host_solver_disarm();
bool dmsa_host_lu_solve(const std::vector<double>& A, int n, const double* b, std::vector<double>& x);
bool dmsa_host_chol_solve(const std::vector<double>& A, int n, const double* b, std::vector<double>& x);
namespace {
inline bool lu_solve_vec(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) { return dmsa_host_lu_solve(A, n, b, x); }
inline bool chol_solve_vec(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) { return dmsa_host_chol_solve(A, n, b, x); }

inline int pad32(int v) { return (v + 31) / 32 * 32; }

}  // namespace

enum { MODEL_NONE = 0, MODEL_TRAJ = 1, MODEL_KF = 2 };

struct dmsa_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // second stream: the two resolution levels of a set build run side by side (fork / join with events)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evLevel0 = nullptr;
    std::vector<cudaEvent_t> evScan;  // one per uploaded scan (register_scans)
    std::string err;
    int64_t launches = 0;
    // programmatic dependent launch: nesting depth of the PdlScopes, and per stream (0: stream, 1: stream2) the epoch of its last launch
    int pdlDepth = 0;
    unsigned long long pdlEpoch[2] = {0, 0};
    int model = MODEL_NONE;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;  // set by dmsa_b200_comm_init: the row-sharded iteration all-reduces [H | g | e0^T e0] and the 9 trial costs
    int64_t collectives = 0;    // NCCL all-reduces issued on the context's stream
    bool worldValid = false;  // globalPoints on the device correspond to the staged points (set by the base transform)
    int solverMode = 0;  // 0 (default): LM step solved on the device for P <= 128 (kernels_solve.cuh), 1: on the host (host_solve.cpp); bit-identical
    int meanMode = 0;  // 0: order-free exactly-rounded mean (default), 1: the reference's sequential float accumulation

    HostPoses poses;
    double origin[3] = {0, 0, 0};
    float minGridSize = 0.3f;  // OptimizablePointSet.h:24

    // trajectory timing (ContinuousTrajectory.h:301-346)
    double t0 = 0, horizon = 0, dt_res = 1e-3;
    int n_total = 0;
    bool tabValid = false;  // timing tables in HBM correspond to (tabHorizon, tabPoses, tabDt)
    double tabHorizon = 0, tabDt = 0;
    int tabPoses = 0;
    bool useImu = false, imuSet = false;
    std::vector<double> stamps, trajTime;
    std::vector<int> paramIndices;
    DBuf<double> d_stamps, d_trajTime, d_urel, d_fh;
    DBuf<int> d_seg, d_hit, d_paramIdx;
    DBuf<double> d_imu;  // preRot | prePos | preVel | covInv
    double balancingImu = 0.001f, gravity[3] = {0.0, 0.0, -9.805};

    // keyframe factors
    bool useGrav = false, useOdom = false;
    DBuf<double> d_kfD;  // measGrav | odomT | odomR
    DBuf<int> d_plausible;
    double balanceGrav = 1.0, balanceOdom = 1000.0;
    std::vector<std::vector<dmsa_b200_point_normal>> kfClouds;
    std::vector<std::vector<int>> kfRings;
    std::vector<float> kfGrid;

    // points in HBM
    int64_t n_scan = 0, n_static = 0;
    DBuf<unsigned char> d_stage;
    DBuf<float4> d_local, d_world, d_normal_l, d_normal_w;
    DBuf<int> d_tid, d_ring, d_flag;

    // pose batches
    DBuf<double> d_p, d_step, d_batch, d_globO, d_globT, d_quat, d_extra, d_dense;
    DBuf<float> d_Mtab, d_Mpair;
    int pairMode = 1;  // 1: pair-packed cost kernels (FMUL2/FADD2) with the shared-rotation fast path for the translation vectors of the
                       // forward-difference batch, 2: pair-packed without the fast path, 0: scalar kernels; all bit-identical
    int curV = 0, curVld = 0;
    int runAhead = 1;  // dmsa_b200_optimize enqueues loop body i + 1 before reading body i's results (device LM solver only)
    double* raPin = nullptr;  // pinned ring of two read-back blocks
    size_t raCap = 0;
    cudaEvent_t raEv[2] = {nullptr, nullptr};
    bool fdBatch = false;  // the tables hold the forward-difference batch [p, p + h e_0, ..]: vectors beyond 3 (n - 1) perturb translations only

    // set construction
    DBuf<LevelInfo> d_linfo;
    LevelInfo h_linfo[2];
    DBuf<int> d_keys, d_bb, d_idx, d_sidx, d_flagA, d_scanA, d_raw_start, d_raw_diff, d_acc_flag, d_acc_scan, d_out_cnt, d_sub, d_ntile, d_tile_off,
        d_best_ij, d_scratch;
    DBuf<SplitTile> d_tiles;
    DBuf<float> d_best_v;
    DBuf<int> d_split_search, d_split_nbox;
    DBuf<unsigned long long> d_code, d_scode;
    DBuf<unsigned char> d_ctl;  // tickets, digit histograms and look-back status words of the hand-written sort / scans (zeroed per build)
    bool sortAttr = false;
    DBuf<float4> d_rec, d_wrec;
    DBuf<int> d_cell_start, d_cell_n, d_cell_level, d_cell_key, d_cell_sub, d_cell_kind, d_nchunk, d_chunk_off, d_okey, d_oval;
    DBuf<float> d_cell_info, d_cell_w0, d_cell_w;
    DBuf<Chunk> d_chunks;
    int G = 0;       // accepted sets of the last build (host copy; -1 while a deferred build's count is still on the device only)
    int Gb = 0;      // bound the per-set grids / buffers of the last build were sized for (== G after a synchronous build)
    int Gguess = 0;  // set count of the previous build of this context: sizes the grids of a deferred build (verified late)
    int64_t M = 0;
    bool levelOn[2] = {false, false};
    int cachedDepth[2] = {0, 0};
    int cellCap = 0;

    // cost
    DBuf<double> d_mom;  // centred second moments of every set [g][6]
    // uniform grids for the radius queries of addStaticPoints / getOverlap
    DBuf<int> d_gbucket, d_gstart, d_gcursor, d_gcount;
    DBuf<float4> d_gpts, d_gquery;
    DBuf<unsigned char> d_gsel, d_gctl;
    // the window grid of addStaticPoints is kept across the keyframe clouds of one call (DmsaSlam.h:296-339: one kd-tree)
    uint64_t worldEpoch = 0, gridEpoch = ~0ull;
    float gridRadius = -1.0f;
    int gridN = 0, gridB = 0;
    double gridH = 0;
    // buffers of the pre-processing / normal estimation entry points (dmsa_b200_pre.inl)
    DBuf<LevelInfo> p_linfo;
    DBuf<int> p_keys, p_bb, p_idx, p_sidx, p_scan, p_raw_start, p_raw_diff, p_pick, p_flag, p_pos, p_rand, p_nn;
    DBuf<unsigned long long> p_code, p_scode;
    DBuf<unsigned char> p_ctl, p_raw, p_out;
    DBuf<float4> p_pts, p_cloud;
    DBuf<float> p_range;
    DBuf<double> d_S, d_Q, d_E, d_jpart, d_hg, d_ls, d_lspart, d_solve, d_iter, d_chol;
    bool solveAttr = false;
    DBuf<int> d_biglist;  // sets with more than GAUSS_WARP_MAX members (+ the count at [cellCap])
    DBuf<int> d_done;  // per-set completion counters of k_cost_quad [0, cap] and k_cost_sum [cap + 1, 2 cap + 1] (zeroed by the set build, self-resetting)
    DBuf<float> d_mu;  // means of the sets cut into more than MEAN_INLINE_MAX chunks [(g*3 + a) * Vld + v]
    size_t chunkBound = 0;

    // host mirrors; the small per-iteration read-backs / uploads go through one pinned block so that
    // cudaMemcpyAsync is a true async DMA (pageable memory costs a staging copy and an implicit synchronisation)
    std::vector<double> h_hg, h_p;
    double* pin = nullptr;  // [P*P+P+1 hg | 16 ls | P params | P step] doubles, then 2 LevelInfo
    size_t pinCap = 0;
    cudaEvent_t evUpload = nullptr;  // last H2D out of the pinned block (waited on before the block is rewritten)
    double lastErr0 = 0;

    // optional per-kernel timing with CUDA events on the launching stream (bench.py roofline)
    bool profiling = false;
    int phase = 0;  // 0: forward-difference batch, 1: line-search batch
    std::vector<cudaEvent_t> evPool;
    struct Span { int id; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double profMs[32] = {0};
    int64_t profCount[32] = {0};
};

enum {
    PROF_POSE_FD = 0, PROF_POSE_LS, PROF_TRANSFORM, PROF_SETS_KEYS, PROF_SETS_SORT, PROF_SETS_STATS,
    PROF_FUSED_FD, PROF_FUSED_LS, PROF_SUM_FD, PROF_SUM_LS, PROF_MEAN_FD, PROF_MEAN_LS, PROF_QUAD_FD, PROF_QUAD_LS, PROF_FIN_FD, PROF_FIN_LS, PROF_JTJ, PROF_COLSUM, PROF_LM_SOLVE, PROF_HOST_SOLVE, PROF_HOST_ITER,
    PROF_NUM
};
static const char* kProfNames[PROF_NUM] = {
    "pose_tables_fd", "pose_tables_ls", "transform_points", "sets_keys_root", "sets_sort_segment", "sets_gaussians_chunks",
    "k_cost_fused_fd", "k_cost_fused_ls", "k_cost_sum_fd", "k_cost_sum_ls", "k_cost_mean_fd", "k_cost_mean_ls", "k_cost_quad_fd", "k_cost_quad_ls", "k_cost_fin_fd", "k_cost_fin_ls",
    "k_jtj", "k_col_sumsq", "k_lm_solve", "host_lm_solve_wall", "host_iteration_wall"};

static cudaEvent_t profEvent(dmsa_b200_ctx* ctx) {
    cudaEvent_t e;
    if (!ctx->evPool.empty()) {
        e = ctx->evPool.back();
        ctx->evPool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
struct ProfScope {
    dmsa_b200_ctx* ctx;
    int id;
    cudaEvent_t a = nullptr;
    ProfScope(dmsa_b200_ctx* c, int i) : ctx(c), id(i) {
        if (ctx->profiling) {
            a = profEvent(ctx);
            cudaEventRecord(a, ctx->stream);
        }
    }
    ~ProfScope() {
        if (a) {
            cudaEvent_t b = profEvent(ctx);
            cudaEventRecord(b, ctx->stream);
            ctx->spans.push_back({id, a, b});
        }
    }
};
static void profCollect(dmsa_b200_ctx* ctx) {  // call after a stream synchronize
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            ctx->profMs[sp.id] += ms;
            ctx->profCount[sp.id]++;
        }
        ctx->evPool.push_back(sp.a);
        ctx->evPool.push_back(sp.b);
    }
    ctx->spans.clear();
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return DMSA_B200_ERR_CUDA;                                                             \
        }                                                                                          \
    } while (0)
#define CKRC(call)                 \
    do {                           \
        int rc__ = (call);         \
        if (rc__ != 0) return rc__; \
    } while (0)
// Kernels go out through cudaLaunchKernelEx; inside a PdlScope a launch whose predecessor on the stream is a kernel of this
// library carries the programmatic-stream-serialization attribute (pdl.cuh): it is set up while the predecessor still
// runs and its threads wait in DMSA_PDL_ENTER() until that one has completed.  DMSA_B200_PDL=0 in the environment
// launches everything plainly (the kernels' griddepcontrol instructions are then no-ops).
static const bool g_pdl = !(getenv("DMSA_B200_PDL") && atoi(getenv("DMSA_B200_PDL")) == 0);
struct PdlScope {
    dmsa_b200_ctx* c;
    explicit PdlScope(dmsa_b200_ctx* ctx) : c(ctx) {
        if (c->pdlDepth++ == 0) pdlBreak();
    }
    ~PdlScope() { --c->pdlDepth; }
    PdlScope(const PdlScope&) = delete;
    PdlScope& operator=(const PdlScope&) = delete;
};
template <typename... KArgs, typename... Args>
static inline void launchKernel(dmsa_b200_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t strm, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = strm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    const unsigned long long ep = g_opEpoch.load(std::memory_order_relaxed);
    int slot = -1;
    if (strm == ctx->stream) slot = 0;
    else if (strm == ctx->stream2) slot = 1;
    const bool chain = g_pdl && ctx->pdlDepth > 0 && slot >= 0 && ctx->pdlEpoch[slot] == ep && !g_commSeen.load(std::memory_order_relaxed);
    cfg.numAttrs = chain ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // (errors surface at the next cudaGetLastError, like <<< >>>)
    if (slot >= 0) ctx->pdlEpoch[slot] = ep;
}
#define LAUNCH(kern, grid, block, smem, ...)                                             \
    do {                                                                                 \
        launchKernel(ctx, kern, (grid), (block), (smem), ctx->stream, __VA_ARGS__);      \
        ctx->launches++;                                                                 \
    } while (0)
#define LAUNCH_ON(strm, kern, grid, block, smem, ...)                            \
    do {                                                                         \
        launchKernel(ctx, kern, (grid), (block), (smem), (strm), __VA_ARGS__);   \
        ctx->launches++;                                                         \
    } while (0)
#define ARGFAIL(msg)               \
    do {                           \
        ctx->err = (msg);          \
        return DMSA_B200_ERR_ARG;  \
    } while (0)

namespace {

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

__global__ void k_unpack_psi(const unsigned char* __restrict__ raw, int n, int out_off, const double* __restrict__ trajTime, int n_total, double t0,
                             int is_static, float4* __restrict__ local, float4* __restrict__ world, int* __restrict__ ring, int* __restrict__ tid,
                             int* __restrict__ flag) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = reinterpret_cast<const float4*>(raw)[2 * (size_t)i];
    const double stamp = *reinterpret_cast<const double*>(raw + 32 * (size_t)i + 16);
    const int id = *reinterpret_cast<const int*>(raw + 32 * (size_t)i + 24);
    if (p.w != 1.0f) atomicOr(flag, 1);
    if ((unsigned)id > 65535u) atomicOr(flag + 1, 1);  // sticky: the set build then takes its ring ids by gather (kernels_sort.cuh ring_packed)
    const int o = out_off + i;
    local[o] = p;
    ring[o] = id;
    if (is_static) {
        world[o] = p;  // addStaticPoints: ContinuousTrajectory.h:164-171
        tid[o] = -1;
    } else {
        // registerPcBuffer: std::lower_bound(trajTime, stamp - t0), clamped to n_total-1   ContinuousTrajectory.h:253-254
        const double val = stamp - t0;
        int lo = 0, hi = n_total;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (trajTime[mid] < val)
                lo = mid + 1;
            else
                hi = mid;
        }
        tid[o] = min(lo, n_total - 1);
    }
}

__global__ void k_unpack_pn(const unsigned char* __restrict__ raw, const int* __restrict__ rings, int n, int out_off, int kf, float4* __restrict__ local,
                            float4* __restrict__ normal, int* __restrict__ ring, int* __restrict__ tid, int* __restrict__ flag) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* r = reinterpret_cast<const float4*>(raw) + 3 * (size_t)i;
    const float4 p = r[0];
    if (p.w != 1.0f) atomicOr(flag, 1);
    const int o = out_off + i;
    local[o] = p;
    normal[o] = r[1];
    ring[o] = rings[i];
    if ((unsigned)rings[i] > 65535u) atomicOr(flag + 1, 1);
    tid[o] = kf;
}

// centralize / decentralize static points: point -= float(origin) / += float(origin)    ContinuousTrajectory.h:83-87, 95-99
__global__ void k_shift_static(float4* __restrict__ local, float4* __restrict__ world, int off, int n, float ox, float oy, float oz, int sign) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = world[off + i];
    if (sign < 0) {
        p.x = __fsub_rn(p.x, ox);
        p.y = __fsub_rn(p.y, oy);
        p.z = __fsub_rn(p.z, oz);
    } else {
        p.x = __fadd_rn(p.x, ox);
        p.y = __fadd_rn(p.y, oy);
        p.z = __fadd_rn(p.z, oz);
    }
    world[off + i] = p;
    local[off + i] = p;
}

// ---- keyframe-bundle extension: a bundle's [H_b | g_b | err0_b] scattered into the global system (parameters are RELATIVE
// poses, so bundle-local parameter i is global parameter idx[i]); the bundles of a rank run on one stream, one after the
// other, so plain read-modify-write is race-free and the summation order is fixed
__global__ void k_bundle_scatter(const double* __restrict__ hg, int Pl, const int* __restrict__ idx, int Pg, const LevelInfo* __restrict__ li,
                                 int depth0, int depth1, double* __restrict__ ghg) {
    DMSA_PDL_ENTER();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int nH = Pl * Pl;
    if (e < nH) {
        const int i = e / Pl, j = e - i * Pl;
        ghg[(size_t)idx[i] * Pg + idx[j]] += hg[e];
    } else if (e < nH + Pl) {
        ghg[(size_t)Pg * Pg + idx[e - nH]] += hg[e];
    } else if (e == nH + Pl) {
        ghg[(size_t)Pg * Pg + Pg] += hg[e];                           // e0^T e0
        ghg[(size_t)Pg * Pg + Pg + 1] += (double)total_sets(li);      // number of Gaussian sets (DmsaOptimizer.h:89)
        // a deferred build that ran on a wrong guess (more sets than the grids were sized for, or a deeper octree than the
        // sort keys covered) is counted here, so that after the all-reduce EVERY rank knows the iteration must be repeated
        const bool miss = li[0].G + li[1].G > max(li[0].bound, li[1].bound) || li[0].depth > depth0 || li[1].depth > depth1 || li[0].error || li[1].error;
        if (miss) ghg[(size_t)Pg * Pg + Pg + 2] += 1.0;
    }
}
__global__ void k_bundle_gather_step(const double* __restrict__ gstep, const int* __restrict__ idx, int Pl, double* __restrict__ step) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Pl) step[i] = gstep[idx[i]];
}
// The tail of one loop body on the device (DmsaOptimizer.h:113-143, 152-182): line-search winner, parameter update, stop tests.
// rec = the iteration's read-back block: [0, 9) trial costs | 9 best k | 10 stop code | 11 ||step|| | [16, 16 + P) step |
// 16 + P err0 | 16 + P + 1 NaN flag | then P new parameters | P parameters of the last trial point.  d_p is advanced in
// place, so the next loop body can be enqueued before the host has seen this one.  The expressions are the host's
// (iterationImpl), operation for operation.
__global__ void k_iter_decide(double* __restrict__ rec, double* __restrict__ p, int P, double epsilon) {
    DMSA_PDL_ENTER();
    __shared__ int s_best;
    const double* step = rec + 16;
    const double err0 = rec[16 + P];
    const bool nan = rec[16 + P + 1] != 0.0;  // (2.0 = the Cholesky solver refused the system: parameters stay, the host redoes the body)
    if (threadIdx.x == 0) {
        double minError = err0;
        int best = 0;
        for (int k = 1; k < 10; ++k)
            if (rec[k - 1] < minError) {
                minError = rec[k - 1];
                best = k;
            }
        double nrm = 0;
        for (int i = 0; i < P; ++i) nrm += step[i] * step[i];
        nrm = sqrt(nrm);
        int stop = DMSA_B200_STOP_MAX_ITER;
        if (nan)
            stop = DMSA_B200_STOP_NAN;
        else if (best == 0)
            stop = DMSA_B200_STOP_NO_IMPROVEMENT;
        else if (nrm < epsilon)
            stop = DMSA_B200_STOP_EPSILON;
        s_best = best;
        rec[9] = (double)best;
        rec[10] = (double)stop;
        rec[11] = nrm;
    }
    __syncthreads();
    const int best = s_best;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const double pv = p[i];
        const double plast = pv + 0.1 * 9.0 * step[i];
        double pnew = pv;
        if (!nan) pnew = best == 0 ? plast : pv + 0.1 * (double)best * step[i];
        rec[16 + P + 2 + i] = pnew;
        rec[16 + 2 * P + 2 + i] = plast;
        p[i] = pnew;
    }
}
__global__ void k_add9(const double* __restrict__ src, double* __restrict__ dst) {
    DMSA_PDL_ENTER();
    if (threadIdx.x < 9) dst[threadIdx.x] += src[threadIdx.x];
}

int64_t numPoints(const dmsa_b200_ctx* ctx) { return ctx->n_scan + ctx->n_static; }
int numExtra(const dmsa_b200_ctx* ctx) {
    if (ctx->rank != 0) return 0;
    if (ctx->model == MODEL_TRAJ) return (ctx->useImu && ctx->imuSet) ? ctx->poses.n - 1 : 0;
    if (ctx->model == MODEL_KF) return (ctx->useGrav ? ctx->poses.n : 0) + (ctx->useOdom ? ctx->poses.n - 1 : 0);
    return 0;
}
int numTableRows(const dmsa_b200_ctx* ctx) { return ctx->model == MODEL_TRAJ ? ctx->n_total : ctx->poses.n; }

// pose chain (+ dense table) for the V vectors currently in d_batch
int runPoseTables(dmsa_b200_ctx* ctx, int V) {
    PdlScope pdl_(ctx);
    const int P = 6 * (ctx->poses.n - 1), n = ctx->poses.n, Vld = pad32(V), rows = numTableRows(ctx);
    const int E = numExtra(ctx);
    CK(ctx->d_globO.ensure((size_t)3 * n * Vld));
    CK(ctx->d_globT.ensure((size_t)3 * n * Vld));
    CK(ctx->d_quat.ensure((size_t)4 * n * Vld));
    CK(ctx->d_Mtab.ensure((size_t)(rows + 1) * Vld * 12));  // + the identity row used by static points
    const bool pairTab = ctx->pairMode && V > 16 && ctx->meanMode == 0;
    if (pairTab) CK(ctx->d_Mpair.ensure((size_t)(rows + 1) * Vld * 12));
    if (E > 0) CK(ctx->d_extra.ensure((size_t)E * Vld));
    PoseBatch pb;
    pb.V = V;
    pb.Vld = Vld;
    pb.P = P;
    pb.n = n;
    pb.params = ctx->d_batch.p;
    for (int a = 0; a < 3; ++a) {
        pb.pose0[a] = ctx->poses.relO[a];
        pb.pose0[3 + a] = ctx->poses.relT[a];
    }
    pb.globO_t = ctx->d_globO.p;
    pb.globT_t = ctx->d_globT.p;
    pb.quat_t = ctx->d_quat.p;
    pb.extra = E > 0 ? ctx->d_extra.p : nullptr;
    pb.Mpair = pairTab ? ctx->d_Mpair.p : nullptr;
    TrajTiming tt;
    tt.n_total = ctx->n_total;
    tt.seg = ctx->d_seg.p;
    tt.urel = ctx->d_urel.p;
    tt.fh = ctx->d_fh.p;
    tt.hit = ctx->d_hit.p;
    ImuFactors imu;
    memset(&imu, 0, sizeof(imu));
    KfFactors kf;
    memset(&kf, 0, sizeof(kf));
    const size_t smem = (size_t)n * 24 * sizeof(double);
    ctx->fdBatch = false;
    ProfScope prof_(ctx, ctx->phase ? PROF_POSE_LS : PROF_POSE_FD);
    if (ctx->model == MODEL_TRAJ) {
        if (E > 0) {
            imu.enabled = 1;
            imu.preRot = ctx->d_imu.p;
            imu.prePos = imu.preRot + 9 * (size_t)n;
            imu.preVel = imu.prePos + 3 * (size_t)n;
            imu.covInv = imu.preVel + 3 * (size_t)n;
            imu.paramIdx = ctx->d_paramIdx.p;
            imu.stamps = ctx->d_stamps.p;
            imu.balancing = ctx->balancingImu;
            imu.dt_res = ctx->dt_res;
            for (int a = 0; a < 3; ++a) imu.gravity[a] = ctx->gravity[a];
        }
        LAUNCH(k_pose_chain<0>, V, 64, smem, pb, ctx->d_Mtab.p, imu, kf, tt);
        dim3 blk(32, 4), grd(Vld / 32, cdiv(ctx->n_total + 1, 4));
        LAUNCH(k_dense_table, grd, blk, 0, pb, tt, ctx->d_Mtab.p);
    } else {
        kf.useGrav = ctx->useGrav && E > 0;
        kf.useOdom = ctx->useOdom && E > 0;
        kf.measGrav = ctx->d_kfD.p;
        kf.odomT = kf.measGrav ? kf.measGrav + 3 * (size_t)n : nullptr;
        kf.odomR = kf.odomT ? kf.odomT + 3 * (size_t)n : nullptr;
        kf.plausible = ctx->d_plausible.p;
        kf.balanceGrav = ctx->balanceGrav;
        kf.balanceOdom = ctx->balanceOdom;
        for (int a = 0; a < 3; ++a) kf.gravity[a] = ctx->gravity[a];
        LAUNCH(k_pose_chain<1>, V, 64, smem, pb, ctx->d_Mtab.p, imu, kf, tt);
    }
    ctx->curV = V;
    ctx->curVld = Vld;
    CK(cudaGetLastError());
    return 0;
}

int ensurePinned(dmsa_b200_ctx* ctx, int P) {
    if (!ctx->evUpload) CK(cudaEventCreateWithFlags(&ctx->evUpload, cudaEventDisableTiming));
    const size_t need = ((size_t)P * P + P + 1 + 16 + 2 * (size_t)P) * sizeof(double) + 2 * sizeof(LevelInfo) + 64;
    if (need <= ctx->pinCap) return 0;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr;
    ctx->pinCap = 0;
    CK(cudaHostAlloc((void**)&ctx->pin, need, cudaHostAllocDefault));
    ctx->pinCap = need;
    return 0;
}
inline double* pinHg(dmsa_b200_ctx* ctx) { return ctx->pin; }
inline double* pinLs(dmsa_b200_ctx* ctx, int P) { return ctx->pin + (size_t)P * P + P + 1; }
inline double* pinParams(dmsa_b200_ctx* ctx, int P) { return pinLs(ctx, P) + 16; }
inline double* pinStep(dmsa_b200_ctx* ctx, int P) { return pinParams(ctx, P) + P; }
inline LevelInfo* pinLinfo(dmsa_b200_ctx* ctx, int P) { return reinterpret_cast<LevelInfo*>(pinStep(ctx, P) + P); }

int uploadParams(dmsa_b200_ctx* ctx) {
    ctx->poses.getParams(ctx->h_p);
    const int P = (int)ctx->h_p.size();
    CKRC(ensurePinned(ctx, P));
    CK(ctx->d_p.ensure(std::max(P, 1)));
    CK(ctx->d_step.ensure(std::max(P, 1)));
    CK(ctx->d_batch.ensure((size_t)(P + 1) * std::max(P, 1)));
    if (P > 0) {
        CK(cudaEventSynchronize(ctx->evUpload));
        std::copy(ctx->h_p.begin(), ctx->h_p.end(), pinParams(ctx, P));
        CK(cudaMemcpyAsync(ctx->d_p.p, pinParams(ctx, P), P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->evUpload, ctx->stream));
    }
    return 0;
}

int transformBase(dmsa_b200_ctx* ctx) {
    PdlScope pdl_(ctx);
    const int64_t N = numPoints(ctx);
    if (N == 0) return 0;
    CK(ctx->d_world.ensure(N));
    const bool kfm = ctx->model == MODEL_KF;
    if (kfm) CK(ctx->d_normal_w.ensure(N));
    // static points (tid = -1) pass through: their world copy is maintained by centralize/decentralize
    ProfScope prof_(ctx, PROF_TRANSFORM);
    LAUNCH(k_transform_points, cdiv(ctx->n_scan, 256), 256, 0, ctx->d_local.p, ctx->d_tid.p, (int)ctx->n_scan, reinterpret_cast<const float4*>(ctx->d_Mtab.p),
           ctx->curVld, 0, ctx->d_world.p, kfm ? ctx->d_normal_l.p : nullptr, kfm ? ctx->d_normal_w.p : nullptr);
    CK(cudaGetLastError());
    ctx->worldValid = true;
    ctx->worldEpoch++;
    return 0;
}

// control block of the hand-written sort / chained scans of one set build (kernels_sort.cuh); one memset zeroes all of it
struct CtlLayout {
    size_t tickets = 0, gbready = 0, bigCount = 0, orderHist = 0, hist = 0, look = 0, segStatus = 0, emitStatus = 0, scanStatus[3] = {0, 0, 0}, bytes = 0;
    int tilesSort = 0, tilesScan = 0;
    static size_t al(size_t x) { return (x + 15) / 16 * 16; }
    CtlLayout(int N, int npass, int cells) {
        tilesSort = (N + RS_TILE - 1) / RS_TILE;
        tilesScan = (std::max(N, cells) + 1 + CS_TILE - 1) / CS_TILE;
        const int tilesEmit = (N + EM_TILE - 1) / EM_TILE + 1;
        size_t o = 0;
        tickets = o;  // 16 ints: [0, 8) sort passes, 8 segment, 9 emit, 10..12 scans
        o = al(o + 16 * sizeof(int));
        gbready = o;
        o = al(o + RS_MAXSEG * sizeof(int));
        bigCount = o;  // number of sets with more than GAUSS_WARP_MAX members (k_gauss_list)
        o = al(o + sizeof(int));
        orderHist = o;  // size-class histogram + cursors of the fused kernel's issue order (k_cell_plan / k_cell_order)
        o = al(o + 2 * ORDER_CLASSES * sizeof(int));
        hist = o;
        o = al(o + (size_t)RS_MAXSEG * RS_MAXPASS * RS_BINS * sizeof(u32_t));
        look = o;
        o = al(o + (size_t)npass * RS_MAXSEG * tilesSort * RS_BINS * sizeof(u32_t));
        segStatus = o;
        o = al(o + (size_t)RS_MAXSEG * tilesScan * sizeof(u64_t));
        emitStatus = o;
        o = al(o + (size_t)RS_MAXSEG * tilesEmit * sizeof(u64_t));
        for (int k = 0; k < 3; ++k) {
            scanStatus[k] = o;
            o = al(o + (size_t)tilesScan * sizeof(u64_t));
        }
        bytes = o;
    }
};
// out[i] = in[0] + .. + in[i - 1], i < n (single-pass chained scan; slot selects the ticket / status area of the control block)
int scanExclusive(dmsa_b200_ctx* ctx, cudaStream_t strm, const CtlLayout& cl, int slot, const int* in, int* out, int n) {
    ScanArgs sa;
    sa.in = in;
    sa.out = out;
    sa.n = n;
    sa.tiles = (n + CS_TILE - 1) / CS_TILE;
    sa.status = reinterpret_cast<u64_t*>(ctx->d_ctl.p + cl.scanStatus[slot]);
    sa.ticket = reinterpret_cast<int*>(ctx->d_ctl.p + cl.tickets) + 10 + slot;
    if (sa.tiles > cl.tilesScan) ARGFAIL("scanExclusive: more tiles than the control block was sized for");
    LAUNCH_ON(strm, k_scan_excl, sa.tiles, CS_T, 0, sa);
    return 0;
}

// reset + both createGaussianSets + updateRebalancingWeights + work decomposition   DmsaOptimizer.h:78-96
// defer = true (inside an iteration, when a previous build gives a grid-size guess): no host synchronisation at all; the
// set count stays on the device (LevelInfo::G), the kernels behind the build read it there, and the caller verifies the
// guess (and the octree depth guess) with the iteration's single read-back.
int buildSets(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, bool defer = false) {
    PdlScope pdl_(ctx);
    const int64_t N64 = numPoints(ctx);
    if (N64 <= 0 || N64 > 0x3fffffff) ARGFAIL("build_sets: no points staged (or more than 2^30)");
    if (!ctx->worldValid) ARGFAIL("build_sets: call update_global_points first (the sets are built on globalPoints)");
    ctx->G = 0;
    // splitSet is specialised for PointCloud<PointNormal> only (Gaussians.h:19-28): the trajectory model never splits
    const bool split = st->gauss_split && ctx->model == MODEL_KF;
    const int N = (int)N64;
    const int Ppin = std::max(0, 6 * (ctx->poses.n - 1));
    const int nb = (N + DMSA_KEYS_BLOCK - 1) / DMSA_KEYS_BLOCK;
    const int minPts = st->min_num_points_per_set;
    const int cap = (int)std::min<int64_t>(2 * N64 / std::max(1, minPts) + 16, 2 * N64 + 16);
    ctx->cellCap = cap;
    CK(ctx->d_linfo.ensure(2));
    CK(ctx->d_keys.ensure((size_t)6 * N));
    CK(ctx->d_bb.ensure((size_t)2 * 12 * nb));
    // per-level scratch (x2: level 0 on the context's stream, level 1 on stream2, side by side)
    const size_t N2 = (size_t)N + 2;
    CK(ctx->d_code.ensure((size_t)2 * N));
    CK(ctx->d_scode.ensure((size_t)2 * N));
    CK(ctx->d_idx.ensure((size_t)2 * N));
    CK(ctx->d_sidx.ensure((size_t)2 * N));
    CK(ctx->d_flagA.ensure((size_t)2 * N));
    CK(ctx->d_scanA.ensure((size_t)2 * N));
    CK(ctx->d_raw_start.ensure(2 * N2));
    CK(ctx->d_raw_diff.ensure(2 * N2));
    CK(ctx->d_acc_flag.ensure(2 * N2));
    CK(ctx->d_acc_scan.ensure(2 * N2));
    CK(ctx->d_out_cnt.ensure(2 * N2));
    CK(ctx->d_sub.ensure(2 * 6 * N2));
    const size_t tileBound = (size_t)N / 256 + (size_t)N / std::max(1, minPts) + 2;
    if (split) {
        CK(ctx->d_ntile.ensure(2 * N2));
        CK(ctx->d_tile_off.ensure(2 * N2));
        CK(ctx->d_tiles.ensure(2 * tileBound));
        CK(ctx->d_best_v.ensure(2 * tileBound));
        CK(ctx->d_best_ij.ensure(4 * tileBound));
        CK(ctx->d_scratch.ensure((size_t)2 * N));
        CK(ctx->d_split_nbox.ensure(12 * N2));
        CK(ctx->d_split_search.ensure(2 * N2));
    }
    CK(ctx->d_rec.ensure((size_t)2 * N));
    CK(ctx->d_wrec.ensure((size_t)2 * N));
    CK(ctx->d_cell_start.ensure(cap));
    CK(ctx->d_cell_n.ensure(cap));
    CK(ctx->d_cell_level.ensure(cap));
    CK(ctx->d_cell_key.ensure((size_t)3 * cap));
    CK(ctx->d_cell_sub.ensure(cap));
    CK(ctx->d_cell_kind.ensure(cap));
    CK(ctx->d_cell_info.ensure((size_t)9 * cap));
    CK(ctx->d_cell_w0.ensure(cap));
    CK(ctx->d_cell_w.ensure(cap));
    CK(ctx->d_nchunk.ensure((size_t)cap + 1));
    CK(ctx->d_chunk_off.ensure((size_t)cap + 1));
    if (4 * ((size_t)cap + 1) > ctx->d_done.cap) {
        CK(ctx->d_done.ensure(4 * ((size_t)cap + 1)));
        CK(cudaMemsetAsync(ctx->d_done.p, 0, ctx->d_done.cap * sizeof(int), ctx->stream));
    }
    CellStore cs;
    cs.start = ctx->d_cell_start.p;
    cs.n = ctx->d_cell_n.p;
    cs.level = ctx->d_cell_level.p;
    cs.key = ctx->d_cell_key.p;
    cs.sub = ctx->d_cell_sub.p;
    cs.info = ctx->d_cell_info.p;
    cs.w0 = ctx->d_cell_w0.p;
    cs.w = ctx->d_cell_w.p;

    // a deferred build sizes every per-set grid and row buffer from the previous build's count (verified by the caller)
    for (int l = 0; l < 2; ++l)
        if ((l == 0 ? st->grid_size_1_factor : st->grid_size_2_factor) > std::numeric_limits<float>::min() && ctx->cachedDepth[l] <= 0) defer = false;
    if (ctx->Gguess <= 0) defer = false;
    const int deferBound = defer ? (int)std::min<int64_t>(cap, (int64_t)ctx->Gguess + ctx->Gguess / 4 + 1024) : cap;
    const float factors[2] = {st->grid_size_1_factor, st->grid_size_2_factor};
    for (int l = 0; l < 2; ++l) ctx->levelOn[l] = factors[l] > std::numeric_limits<float>::min();  // DmsaOptimizer.h:81,85
    if (!ctx->levelOn[0] && !ctx->levelOn[1]) CK(cudaMemsetAsync(ctx->d_linfo.p, 0, 2 * sizeof(LevelInfo), ctx->stream));  // (otherwise k_anchor clears the records)
    int prezeroedPasses = -1;  // >= 0: k_keys cleared phase 2's control block for that number of digit passes
    // phase 1: anchors, keys, octree roots
    {
    ProfScope prof_(ctx, PROF_SETS_KEYS);
    LevelPlan plan;
    plan.n = 0;
    for (int l = 0; l < 2; ++l) {
        if (!ctx->levelOn[l]) continue;
        plan.level[plan.n] = l;
        plan.res[plan.n] = factors[l] * ctx->minGridSize;  // float product, DmsaOptimizer.h:82,86
        plan.n++;
    }
    if (plan.n > 0) {
        // with a depth guess the layout of phase 2's control block is known already: k_keys clears it (and the ring-test flags)
        // on the side, and phase 2 starts without its two memsets
        ZeroRanges zr{};
        int npass0 = 1;
        bool guess0 = true;
        for (int l = 0; l < 2; ++l) {
            if (!ctx->levelOn[l]) continue;
            if (ctx->cachedDepth[l] <= 0) guess0 = false;
            npass0 = std::max(npass0, (std::min(64, 3 * ctx->cachedDepth[l] + 1) + 7) / 8);
        }
        if (guess0) {
            const CtlLayout cl0(N, npass0, cap);
            CK(ctx->d_ctl.ensure(cl0.bytes));
            zr.p[0] = reinterpret_cast<unsigned int*>(ctx->d_ctl.p);
            zr.words[0] = cl0.bytes / sizeof(unsigned int);
            zr.p[1] = reinterpret_cast<unsigned int*>(ctx->d_raw_diff.p);
            zr.words[1] = 2 * (size_t)N2;
            prezeroedPasses = npass0;
        }
        LAUNCH(k_anchor, 1, 32, 0, ctx->d_world.p, N, plan, ctx->d_linfo.p, deferBound);
        LAUNCH(k_keys, dim3(nb, plan.n), DMSA_KEYS_BLOCK, 0, ctx->d_world.p, N, plan, ctx->d_linfo.p, ctx->d_keys.p, ctx->d_bb.p, nb, zr);
        LAUNCH(k_root, plan.n, 1024, 0, ctx->d_world.p, N, plan, ctx->d_linfo.p, ctx->d_bb.p, nb);
    }
    }
    // The radix sort needs the number of key bits (3 * octree depth + 1) on the host.  The depth of the previous build of
    // this context is a safe guess (more bits than needed are harmless); it is verified after phase 2 and the phase is
    // redone in the rare case the tree grew.  Without a guess: one synchronisation here.

    int depthUsed[2] = {ctx->cachedDepth[0], ctx->cachedDepth[1]};
    bool haveGuess = true;
    for (int l = 0; l < 2; ++l)
        if (ctx->levelOn[l] && depthUsed[l] <= 0) haveGuess = false;
    if (!haveGuess || ctx->Gguess <= 0) defer = false;
    if (!haveGuess) {
        CKRC(ensurePinned(ctx, Ppin));
        CK(cudaMemcpyAsync(pinLinfo(ctx, Ppin), ctx->d_linfo.p, 2 * sizeof(LevelInfo), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_linfo, pinLinfo(ctx, Ppin), 2 * sizeof(LevelInfo));
        for (int l = 0; l < 2; ++l) {
            if (ctx->levelOn[l] && ctx->h_linfo[l].error) ARGFAIL("build_sets: octree deeper than 21 levels (extent / resolution too large)");
            depthUsed[l] = ctx->h_linfo[l].depth;
        }
    }
phase2:
    // phase 2: sort, segment, accept, emit, gather — both resolution levels in the SAME launches (kernels_sort.cuh): one
    // Morton + histogram kernel, one kernel per 8-bit digit, one chained scan for run heads / leaf starts / ring test, one for
    // acceptance / set numbering / emission (level 1's sets follow level 0's: one scan chain over both levels), one gather.

    int npass = 1;
    for (int l = 0; l < 2; ++l)
        if (ctx->levelOn[l]) npass = std::max(npass, (std::min(64, 3 * depthUsed[l] + 1) + 7) / 8);
    const CtlLayout cl(N, npass, cap);
    CK(ctx->d_ctl.ensure(cl.bytes));
    {
    ProfScope prof_(ctx, PROF_SETS_SORT);
    cudaStream_t strm = ctx->stream;
    if (!ctx->sortAttr) {
        CK(cudaFuncSetAttribute(k_sort_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
        ctx->sortAttr = true;
    }
    if (prezeroedPasses != npass) {  // (no depth guess, or phase 2 is being redone with a deeper tree)
        CK(cudaMemsetAsync(ctx->d_ctl.p, 0, cl.bytes, strm));
        CK(cudaMemsetAsync(ctx->d_raw_diff.p, 0, 2 * N2 * sizeof(int), strm));
    }
    prezeroedPasses = -1;  // a second pass through phase 2 finds the control block used
    SortArgs sa;
    memset(&sa, 0, sizeof(sa));
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    SegmentArgs ga;
    memset(&ga, 0, sizeof(ga));
    EmitArgs ea;
    memset(&ea, 0, sizeof(ea));
    GatherArgs gt;
    memset(&gt, 0, sizeof(gt));
    int nseg = 0;
    // pass p reads A / writes B for even p: the buffers are assigned so that the last pass always lands in d_scode / d_sidx
    const bool odd = (npass & 1) != 0;
    for (int l = 0; l < 2; ++l) {
        if (!ctx->levelOn[l]) continue;
        const int s_ = nseg++;
        u64_t* const code0 = ctx->d_code.p + (size_t)N * l;
        u64_t* const code1 = ctx->d_scode.p + (size_t)N * l;
        u32_t* const idx0 = reinterpret_cast<u32_t*>(ctx->d_idx.p + (size_t)N * l);
        u32_t* const idx1 = reinterpret_cast<u32_t*>(ctx->d_sidx.p + (size_t)N * l);
        sa.seg[s_].keyA = odd ? code0 : code1;
        sa.seg[s_].keyB = odd ? code1 : code0;
        sa.seg[s_].valA = odd ? idx0 : idx1;
        sa.seg[s_].valB = odd ? idx1 : idx0;
        pa.level[s_] = l;
        const u64_t* scode = code1;
        u32_t* sidx = idx1;
        ga.code[s_] = scode;
        ga.idx[s_] = sidx;
        ga.scan[s_] = ctx->d_scanA.p + (size_t)N * l;
        ga.raw_start[s_] = ctx->d_raw_start.p + N2 * l;
        ga.raw_diff[s_] = ctx->d_raw_diff.p + N2 * l;
        ga.level[s_] = l;
        ea.raw_start[s_] = ga.raw_start[s_];
        ea.raw_diff[s_] = ga.raw_diff[s_];
        ea.out_cnt[s_] = ctx->d_out_cnt.p + N2 * l;
        ea.sub_start[s_] = ctx->d_sub.p + 6 * N2 * l;
        ea.sub_n[s_] = ea.sub_start[s_] + 2 * N2;
        ea.sub_code[s_] = ea.sub_n[s_] + 2 * N2;
        ea.idx[s_] = sidx;
        ea.keys[s_] = ctx->d_keys.p + (size_t)3 * N * l;
        ea.mbase[s_] = l * N;
        ea.level[s_] = l;
        gt.idx[s_] = sidx;
        gt.rec[s_] = ctx->d_rec.p + (size_t)N * l;
        gt.wrec[s_] = ctx->d_wrec.p + (size_t)N * l;
        gt.level[s_] = l;
    }
    if (nseg > 0) {
        sa.nseg = nseg;
        sa.n = N;
        sa.tiles = cl.tilesSort;
        sa.npass = npass;
        sa.iota = 1;
        sa.hist = reinterpret_cast<u32_t*>(ctx->d_ctl.p + cl.hist);
        sa.look = reinterpret_cast<u32_t*>(ctx->d_ctl.p + cl.look);
        sa.ticket = reinterpret_cast<int*>(ctx->d_ctl.p + cl.tickets);
        pa.keys_all = ctx->d_keys.p;
        pa.infos = ctx->d_linfo.p;
        pa.ring = ctx->d_ring.p;
        pa.flags = ctx->d_flag.p;
        LAUNCH_ON(strm, k_sort_prepare, dim3(cl.tilesSort, nseg), RS_T, 0, sa, pa);
        for (int p_ = 0; p_ < npass; ++p_) {
            sa.pass = p_;
            LAUNCH_ON(strm, k_sort_pass, cl.tilesSort * nseg, RS_T, RS_SMEM, sa);
        }
        const int tilesSeg = (N + SG_TILE - 1) / SG_TILE, tilesEmit = (N + EM_TILE - 1) / EM_TILE;
        ga.infos = ctx->d_linfo.p;
        ga.ring = ctx->d_ring.p;
        ga.flags = ctx->d_flag.p;
        ga.n = N;
        ga.tiles = tilesSeg;
        ga.status = reinterpret_cast<u64_t*>(ctx->d_ctl.p + cl.segStatus);
        ga.ticket = reinterpret_cast<int*>(ctx->d_ctl.p + cl.tickets) + 8;
        LAUNCH_ON(strm, k_segment, tilesSeg * nseg, SG_T, 0, ga);
        ea.infos = ctx->d_linfo.p;
        ea.nseg = nseg;
        ea.n = N;
        ea.tiles = tilesEmit;
        ea.minPts = minPts;
        ea.cap = cap;
        ea.cs = cs;
        ea.status = reinterpret_cast<u64_t*>(ctx->d_ctl.p + cl.emitStatus);
        ea.ticket = reinterpret_cast<int*>(ctx->d_ctl.p + cl.tickets) + 9;
        ea.gbase_ready = reinterpret_cast<int*>(ctx->d_ctl.p + cl.gbready);
        if (split) {  // Gaussians.h:27-85 on every accepted leaf, then emission from the plan
            for (int s_ = 0; s_ < nseg; ++s_) {
                const int l = ea.level[s_];
                LevelInfo* li = ctx->d_linfo.p + l;
                int* sidx = reinterpret_cast<int*>(const_cast<u32_t*>(ea.idx[s_]));
                int* scanA = ga.scan[s_];
                int* raw_start = ga.raw_start[s_];
                int* acc_flag = ctx->d_acc_flag.p + N2 * l;
                int* out_cnt = ctx->d_out_cnt.p + N2 * l;
                int* sub_start = ctx->d_sub.p + 6 * N2 * l;
                int* sub_n = sub_start + 2 * N2;
                int* sub_code = sub_n + 2 * N2;
                int* ntile = ctx->d_ntile.p + N2 * l;
                int* tile_off = ctx->d_tile_off.p + N2 * l;
                SplitTile* tiles = ctx->d_tiles.p + tileBound * l;
                float* best_v = ctx->d_best_v.p + tileBound * l;
                int* best_i = ctx->d_best_ij.p + 2 * tileBound * l;
                int* best_j = best_i + tileBound;
                int* scratch = ctx->d_scratch.p + (size_t)N * l;
                int* nbox = ctx->d_split_nbox.p + 6 * N2 * l;
                int* search = ctx->d_split_search.p + N2 * l;
                LAUNCH_ON(strm, k_accept, cdiv(N, 256), 256, 0, raw_start, ga.raw_diff[s_], li, minPts, acc_flag, out_cnt, sub_start, sub_n, sub_code);
                LAUNCH_ON(strm, k_split_nbox_init, cdiv(N2, 256), 256, 0, nbox, search, (int)N2, li);
                LAUNCH_ON(strm, k_split_nbox, cdiv(N, 256), 256, 0, sidx, scanA, acc_flag, li, ctx->d_normal_w.p, nbox);
                LAUNCH_ON(strm, k_split_prefilter, cdiv(N, 256), 256, 0, sidx, scanA, acc_flag, li, ctx->d_normal_w.p, nbox, search);
                LAUNCH_ON(strm, k_split_tile_counts, cdiv((size_t)N + 1, 256), 256, 0, raw_start, acc_flag, search, li, ntile);
                CKRC(scanExclusive(ctx, strm, cl, 1 + s_, ntile, tile_off, N + 1));
                LAUNCH_ON(strm, k_split_tile_fill, cdiv(N, 256), 256, 0, ntile, tile_off, li, tiles);
                LAUNCH_ON(strm, k_split_pairs, (unsigned)std::min<size_t>(tileBound, 148 * 8), 256, 0, tiles, tile_off, li, raw_start, sidx, ctx->d_normal_w.p, best_v, best_i, best_j);
                LAUNCH_ON(strm, k_split_decide, 148 * 8, 256, 0, acc_flag, ntile, tile_off, li, raw_start, sidx, scratch, ctx->d_normal_w.p, ctx->d_ring.p, best_v,
                          best_i, best_j, minPts, out_cnt, sub_start, sub_n, sub_code);
            }
            LAUNCH_ON(strm, k_emit<true>, std::min(tilesEmit * nseg, 148 * 4), EM_T, 0, ea);
        } else {
            LAUNCH_ON(strm, k_emit<false>, std::min(tilesEmit * nseg, 148 * 4), EM_T, 0, ea);
        }
        gt.infos = ctx->d_linfo.p;
        LAUNCH_ON(strm, k_gather2, dim3(cdiv(N, 256), nseg), 256, 0, gt, ctx->d_local.p, ctx->d_tid.p, numTableRows(ctx), ctx->d_world.p);
    }
    }
    int Gb = 0;
    if (defer) {
        Gb = deferBound;
        ctx->G = -1;
    } else {
    CKRC(ensurePinned(ctx, Ppin));
    CK(cudaMemcpyAsync(pinLinfo(ctx, Ppin), ctx->d_linfo.p, 2 * sizeof(LevelInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_linfo, pinLinfo(ctx, Ppin), 2 * sizeof(LevelInfo));
    CK(cudaGetLastError());
    {
        bool redo = false;
        for (int l = 0; l < 2; ++l) {
            if (!ctx->levelOn[l]) continue;
            if (ctx->h_linfo[l].error) ARGFAIL("build_sets: octree deeper than 21 levels (extent / resolution too large)");
            if (ctx->h_linfo[l].depth > depthUsed[l]) redo = true;
            depthUsed[l] = ctx->h_linfo[l].depth;
            ctx->cachedDepth[l] = ctx->h_linfo[l].depth;
        }
        if (redo) {
            goto phase2;
        }
    }
    int G = 0;
    for (int l = 0; l < 2; ++l)
        if (ctx->levelOn[l]) G += ctx->h_linfo[l].G;
    if (G > cap) ARGFAIL("build_sets: Gaussian store capacity exceeded");
    ctx->G = G;
    ctx->Gguess = G;
    ctx->Gb = G;
    ctx->M = 0;
    if (G == 0) return 0;
    Gb = G;
    }
    ctx->Gb = Gb;
    const LevelInfo* li = ctx->d_linfo.p;
    // phase 3: per-set statistics, weights, chunk list (grids sized for Gb sets; the kernels read the count on the device)
    ProfScope prof_(ctx, PROF_SETS_STATS);
    CK(ctx->d_okey.ensure((size_t)cap));
    CK(ctx->d_oval.ensure((size_t)2 * cap + 4 * ORDER_CLASSES));
    CK(ctx->d_biglist.ensure((size_t)cap + 1));
    CK(ctx->d_mom.ensure((size_t)6 * cap));
    // stream2: the compact list of the sets with more than GAUSS_WARP_MAX members and their statistics (one block each).
    // Context's stream, side by side: the small sets (one warp each) and the work decomposition of the cost kernels (set
    // kinds, issue order, chunk list), which depends on the set sizes only.  Disjoint outputs; joined before the eigen finish.
    ctx->chunkBound = (size_t)2 * N / CHUNK + (size_t)2 * N / FUSE_MAX + 2;  // big sets only
    CK(ctx->d_chunks.ensure(ctx->chunkBound));
    {
        cudaStream_t s2 = ctx->stream2;
        // (both live in the build's control block: zeroed by its one memset)
        int* cnt = reinterpret_cast<int*>(ctx->d_ctl.p + cl.bigCount);
        int* hist = reinterpret_cast<int*>(ctx->d_ctl.p + cl.orderHist);  // [ORDER_CLASSES histogram | ORDER_CLASSES cursors]
        CK(cudaEventRecord(ctx->evFork, ctx->stream));
        CK(cudaStreamWaitEvent(s2, ctx->evFork, 0));
        LAUNCH_ON(s2, k_gauss_list, cdiv(Gb, 256), 256, 0, cs, li, ctx->d_biglist.p, cnt);
        LAUNCH_ON(s2, k_gaussian_big, 148 * 2, GAUSS_BIG_T, 0, ctx->d_wrec.p, cs, ctx->d_biglist.p, cnt, ctx->d_mom.p);
        LAUNCH_ON(s2, k_weights, 1, 1024, 0, cs, li);  // depends on the set sizes only
        CK(cudaEventRecord(ctx->evJoin, s2));
        LAUNCH(k_gaussian, cdiv((size_t)Gb * 32, 256), 256, 0, ctx->d_wrec.p, cs, li, ctx->d_mom.p);
        // (the completion counters in d_done reset themselves at the end of every cost launch: zeroed when allocated only;
        //  k_cell_plan writes the zeros behind the last set of nchunk itself)
        LAUNCH(k_cell_plan, cdiv((size_t)Gb + 1, 256), 256, 0, cs, li, CHUNK, FUSE_MAX, ctx->rank, ctx->world, Gb, ctx->d_cell_kind.p, ctx->d_nchunk.p, ctx->d_okey.p, hist);
        LAUNCH(k_cell_order, cdiv(Gb, 256), 256, 0, li, ctx->d_okey.p, hist, ctx->d_oval.p + ctx->cellCap + 2 * ORDER_CLASSES);
        CKRC(scanExclusive(ctx, ctx->stream, cl, 0, ctx->d_nchunk.p, ctx->d_chunk_off.p, Gb + 1));  // [Gb] = total
        LAUNCH(k_chunk_fill, cdiv(Gb, 256), 256, 0, cs, li, CHUNK, ctx->d_nchunk.p, ctx->d_chunk_off.p, ctx->d_chunks.p);
        CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
        LAUNCH(k_gaussian_fin, cdiv(Gb, 128), 128, 0, cs, li, ctx->d_mom.p);
    }
    CK(cudaGetLastError());
    return 0;
}

// V cost evaluations on the tables in Mtab -> E rows [0, G) (+ extra rows)
int runCost(dmsa_b200_ctx* ctx) {
    PdlScope pdl_(ctx);
    const int V = ctx->curV, Vld = ctx->curVld, G = ctx->Gb, E = numExtra(ctx);  // G: the bound the build sized its grids for
    if (Vld > 1024) ARGFAIL("more than 1023 pose parameters are not supported by the cost kernels");
    if (ctx->model == MODEL_TRAJ && ctx->useImu && !ctx->imuSet)
        ARGFAIL("cost evaluation: traj_init was called with use_imu = 1 but the IMU factors of this window are missing: call traj_set_imu_factors after traj_init");
    CK(ctx->d_S.ensure(ctx->chunkBound * 3 * Vld));
    CK(ctx->d_Q.ensure(ctx->chunkBound * Vld));
    CK(ctx->d_E.ensure((size_t)(G + E) * Vld));
    CK(ctx->d_mu.ensure((size_t)G * 3 * Vld));
    CostArgs a;
    a.chunks = ctx->d_chunks.p;
    a.n_chunks = ctx->d_chunk_off.p + G;  // the scan ran over Gb + 1 entries, the ones behind the last set are 0
    a.li = ctx->d_linfo.p;
    a.rec = ctx->d_rec.p;
    a.Mtab = reinterpret_cast<const float4*>(ctx->d_Mtab.p);
    a.Mpair = reinterpret_cast<const unsigned long long*>(ctx->d_Mpair.p);
    a.Vp = Vld / 2;
    a.V = V;
    a.Vld = Vld;
    a.S = (V <= 16) ? 32 / V : 1;
    a.order = nullptr;
    a.info = ctx->d_cell_info.p;
    a.w = ctx->d_cell_w.p;
    a.cell_start = ctx->d_cell_start.p;
    a.cell_n = ctx->d_cell_n.p;
    a.cell_kind = ctx->d_cell_kind.p;
    a.nchunk = ctx->d_nchunk.p;
    a.chunk_off = ctx->d_chunk_off.p;
    a.S_part = ctx->d_S.p;
    a.done = ctx->d_done.p;
    a.done1 = ctx->d_done.p + ctx->cellCap + 1;
    a.mu = ctx->d_mu.p;
    a.Q = ctx->d_Q.p;
    a.E = ctx->d_E.p;
    a.order = ctx->d_oval.p + ctx->cellCap + 2 * ORDER_CLASSES;
    const unsigned grid = (unsigned)ctx->chunkBound;
    const int ph = ctx->phase ? 1 : 0;
    if (ctx->meanMode == 1) {
        ProfScope p_(ctx, PROF_FUSED_FD + ph);
        LAUNCH(k_cost_seq, G, Vld, 0, a);
        if (E > 0) LAUNCH(k_append_extra, cdiv((size_t)E * Vld, 256), 256, 0, ctx->d_E.p, ctx->d_linfo.p, ctx->d_extra.p, E, Vld);
        CK(cudaGetLastError());
        return 0;
    }
    const bool packed = a.S > 1;
    const int cls = packed ? 0 : (Vld <= 128 ? 1 : (Vld <= 256 ? 2 : (Vld <= 512 ? 3 : 4)));
#define DISPATCH(KERN, GRID, ...)                                                          \
    switch (cls) {                                                                         \
        case 0: LAUNCH((KERN<true, 32 * PACKED_WARPS, 16>), GRID, 32 * PACKED_WARPS, 0, __VA_ARGS__); break; \
        case 1: LAUNCH((KERN<false, 128, 10>), GRID, Vld, 0, __VA_ARGS__); break;          \
        case 2: LAUNCH((KERN<false, 256, 5>), GRID, Vld, 0, __VA_ARGS__); break;           \
        case 3: LAUNCH((KERN<false, 512, 2>), GRID, Vld, 0, __VA_ARGS__); break;           \
        default: LAUNCH((KERN<false, 1024, 1>), GRID, Vld, 0, __VA_ARGS__); break;         \
    }
    const bool pair = !packed && ctx->pairMode;  // two parameter vectors per thread, packed FP32x2 (kernels_cost.cuh)
    // shared-rotation fast path (kernels_cost.cuh): two blocks per set / chunk, one for the vector pairs that carry a rotation of
    // their own, one for the pairs that perturb a translation only
    int pairT = Vld / 2, mult = 1;
    a.split = 0;
    a.ns = pairT;
    a.done_f = ctx->d_done.p + 2 * ((size_t)ctx->cellCap + 1);
    a.done1_f = ctx->d_done.p + 3 * ((size_t)ctx->cellCap + 1);
    if (pair && ctx->pairMode == 1 && ctx->fdBatch) {
        const int vT = 1 + 3 * (ctx->poses.n - 1);  // first vector that perturbs a translation: p + h e_{3 (n - 1)}
        const int ns = (vT + 1) / 2, nf = (V + 1) / 2 - ns;
        if (nf > 0) {
            a.split = 1;
            a.ns = ns;
            pairT = pad32(std::max(ns, nf));
            mult = 2;
        }
    }
    const int cls2 = pairT <= 32 ? 0 : (pairT <= 64 ? 1 : (pairT <= 128 ? 2 : (pairT <= 256 ? 3 : 4)));
#define DISPATCH2(KERN, GRID, ...)                                                         \
    switch (cls2) {                                                                        \
        case 0: LAUNCH((KERN<32, 24>), (GRID) * mult, pairT, 0, __VA_ARGS__); break;       \
        case 1: LAUNCH((KERN<64, 12>), (GRID) * mult, pairT, 0, __VA_ARGS__); break;       \
        case 2: LAUNCH((KERN<128, 6>), (GRID) * mult, pairT, 0, __VA_ARGS__); break;       \
        case 3: LAUNCH((KERN<256, 3>), (GRID) * mult, pairT, 0, __VA_ARGS__); break;       \
        default: LAUNCH((KERN<512, 1>), (GRID) * mult, pairT, 0, __VA_ARGS__); break;      \
    }
    {
        ProfScope p_(ctx, PROF_FUSED_FD + ph);
        if (pair) {
            DISPATCH2(k_cost_fused2, G, a);
        } else {
            DISPATCH(k_cost_fused, G, a);
        }
    }
    {
        ProfScope p_(ctx, PROF_SUM_FD + ph);
        if (pair) {
            DISPATCH2(k_cost_sum2, grid, a);
        } else {
            DISPATCH(k_cost_sum, grid, a);
        }
    }
    {
        ProfScope p_(ctx, PROF_QUAD_FD + ph);  // includes the per-set mean and the final chunk reduction (last block of a set)
        if (pair) {
            DISPATCH2(k_cost_quad2, grid, a);
        } else {
            DISPATCH(k_cost_quad, grid, a);
        }
    }
#undef DISPATCH2
#undef DISPATCH
    if (E > 0) LAUNCH(k_append_extra, cdiv((size_t)E * Vld, 256), 256, 0, ctx->d_E.p, ctx->d_linfo.p, ctx->d_extra.p, E, Vld);
    CK(cudaGetLastError());
    return 0;
}

// forward-difference batch -> tables (parameters must be uploaded)
int prepareFdBatch(dmsa_b200_ctx* ctx) {
    PdlScope pdl_(ctx);
    const int P = 6 * (ctx->poses.n - 1);
    // `1.0 * sqrt(std::numeric_limits<float>::epsilon())` resolves to sqrt(float)       DmsaOptimizer.h:209
    const double h = 1.0 * (double)sqrtf(FLT_EPSILON);
    LAUNCH(k_make_fd_batch, cdiv((size_t)(P + 1) * P, 256), 256, 0, ctx->d_p.p, P, h, ctx->d_batch.p);
    CKRC(runPoseTables(ctx, P + 1));
    ctx->fdBatch = true;
    return 0;
}

int jtjInto(dmsa_b200_ctx* ctx, double* hg_dev) {
    PdlScope pdl_(ctx);
    const int P = 6 * (ctx->poses.n - 1), E = numExtra(ctx), Vld = ctx->curVld;
    const double h = 1.0 * (double)sqrtf(FLT_EPSILON);
    const double inv_h = 1.0 / h;  // one_div_incr, DmsaOptimizer.h:210
    const int n1 = P + 1;
    const LevelInfo* li = ctx->d_linfo.p;  // the row count G + E is read on the device; so is the row partition (jd_partition)
    if (n1 <= JD_MAXN) {
        // FP64 tensor cores (DMMA m8n8k4): one block per row range keeps every upper-triangular output tile in registers
        CK(ctx->d_jpart.ensure((size_t)JD_MAXBLK * n1 * n1));
        ProfScope prof_(ctx, PROF_JTJ);
        LAUNCH(k_jtj_dmma, JD_MAXBLK, JD_T, 0, ctx->d_E.p, li, E, Vld, P, inv_h, ctx->d_jpart.p);
        LAUNCH(k_jtj_reduce8, cdiv((size_t)n1 * n1, 32), dim3(32, 8), 0, ctx->d_jpart.p, li, E, P, hg_dev);
        CK(cudaGetLastError());
        return 0;
    }
    const int R = ctx->Gb + E;
    int nsplit = std::max(1, std::min(JTJ_MAXSPLIT, (R + 127) / 128));  // upper bound of the device-side split count (monotone in R)
    const int nt = (n1 + JTJ_T - 1) / JTJ_T;
    CK(ctx->d_jpart.ensure((size_t)nsplit * n1 * n1));
    dim3 grid(nt * (nt + 1) / 2, nsplit);
    ProfScope prof_(ctx, PROF_JTJ);
    LAUNCH(k_jtj, grid, 256, 0, ctx->d_E.p, li, E, Vld, P, inv_h, ctx->d_jpart.p);
    LAUNCH(k_jtj_reduce, cdiv((size_t)n1 * n1, 256), 256, 0, ctx->d_jpart.p, li, E, P, hg_dev);
    CK(cudaGetLastError());
    return 0;
}

// the 9 trial costs of adaptiveStepSize for the step in d_step -> ls_dev[9]
int lineSearchDev(dmsa_b200_ctx* ctx, double* ls_dev) {
    PdlScope pdl_(ctx);
    const int P = 6 * (ctx->poses.n - 1);
    LAUNCH(k_make_ls_batch, cdiv((size_t)9 * P, 256), 256, 0, ctx->d_p.p, ctx->d_step.p, P, ctx->d_batch.p);
    ctx->phase = 1;
    int rc = runPoseTables(ctx, 9);
    if (rc == 0) rc = runCost(ctx);
    ctx->phase = 0;
    if (rc) return rc;
    ProfScope prof_(ctx, PROF_COLSUM);
    CK(ctx->d_lspart.ensure(9 * COLSUM_PARTS));
    LAUNCH(k_col_sumsq, dim3(9, COLSUM_PARTS), 256, 0, ctx->d_E.p, ctx->d_linfo.p, numExtra(ctx), ctx->curVld, ctx->d_lspart.p);
    LAUNCH(k_col_sumsq_fin, 1, 32, 0, ctx->d_lspart.p, ls_dev);
    CK(cudaGetLastError());
    return 0;
}
int lineSearchInto(dmsa_b200_ctx* ctx, const double* step_host, double* ls_dev) {
    const int P = 6 * (ctx->poses.n - 1);
    CKRC(ensurePinned(ctx, P));
    CK(cudaEventSynchronize(ctx->evUpload));
    std::copy(step_host, step_host + P, pinStep(ctx, P));
    CK(cudaMemcpyAsync(ctx->d_step.p, pinStep(ctx, P), P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->evUpload, ctx->stream));
    return lineSearchDev(ctx, ls_dev);
}

// LM step on the device (kernels_solve.cuh, P <= 128): hg_dev = [H | g | err0] -> d_step (+ copy in step2), tail = [err0, nan flag]
int lmSolveDev(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, int P, const double* hg_dev, double* step2, double* tail) {
    PdlScope pdl_(ctx);
    if (P > LM_DEV_MAXN) ARGFAIL("lm_solve: the device solver takes at most 128 parameters (larger systems: host solver)");
    const int ld = pad32(P);
    const size_t need = 2 * (size_t)P * ld + P + 2 * (size_t)P + 8;
    if (need > ctx->d_solve.cap) {
        CK(ctx->d_solve.ensure(need));
        CK(cudaMemsetAsync(ctx->d_solve.p, 0, ctx->d_solve.cap * sizeof(double), ctx->stream));
    }
    double* LUT = ctx->d_solve.p;
    double* XT = LUT + (size_t)P * ld;
    double* rdiag = XT + (size_t)P * ld;
    int* piv = reinterpret_cast<int*>(rdiag + P);
    const size_t smem = ((size_t)P * ld + P) * sizeof(double);
    if (!ctx->solveAttr) {
        CK(cudaFuncSetAttribute(k_inv128, cudaFuncAttributeMaxDynamicSharedMemorySize, (LM_DEV_MAXN * LM_DEV_MAXN + LM_DEV_MAXN) * (int)sizeof(double)));
        CK(cudaFuncSetAttribute(k_step_fin, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_DEV_MAXN * LM_DEV_MAXN * (int)sizeof(double)));
        ctx->solveAttr = true;
    }
    ProfScope prof_(ctx, PROF_LM_SOLVE);
    Lu128Args la{hg_dev, P, ld, (double)st->lambda_diag, LUT, rdiag, piv};
    static const bool luDbg = getenv("DMSA_B200_LU_CLK") != nullptr;
    if (luDbg) {
        LAUNCH(k_lu128<true>, 1, LU128_T, 0, la);
        long long c[16];
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpyFromSymbol(c, g_lu_clk, sizeof(c)));
        fprintf(stderr, "k_lu128 step (panel 1, kk 3) cycles: argmax %lld | pivot row staging %lld | bookkeeping+div %lld | panel update %lld ; barrier wait of warp 2 %lld | its trailing update %lld\n",
                c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3], c[6] - c[5], c[7] - c[6]);
        fprintf(stderr, "   trailing steps of warp 2, panel 1: %lld %lld %lld %lld %lld %lld %lld %lld ; barrier exit -> first step %lld\n", c[9] - c[8], c[10] - c[9], c[11] - c[10],
                c[12] - c[11], c[13] - c[12], c[14] - c[13], c[15] - c[14], c[7] - c[15], c[8] - c[6]);
    } else {
        LAUNCH(k_lu128<false>, 1, LU128_T, 0, la);
    }
    Inv128Args ia{LUT, rdiag, piv, P, ld, XT};
    LAUNCH(k_inv128, (P + INV128_WARPS - 1) / INV128_WARPS, INV128_WARPS * 32, smem, ia);
    StepFinArgs fa{hg_dev, XT, P, ld, st->step_length_optim, st->max_step, ctx->d_step.p, step2, tail};
    LAUNCH(k_step_fin, 1, STEPFIN_T, (size_t)P * ld * sizeof(double), fa);
    CK(cudaGetLastError());
    return 0;
}

// LM step by the device Cholesky (kernels_chol.cuh, any P <= 1024): solver mode 2 of the reference-shaped loop and the solver
// of the bundle extension.  step / step2: two copies of the step; tail = [err0, 0 ok / 1 NaN / 2 not positive definite].
int cholSolveDev(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, int n, const double* hg_dev, double* step, double* step2, double* tail) {
    const int ld = pad32(n);
    CK(ctx->d_chol.ensure((size_t)(n + 1) * ld + 8 + ld));
    CholArgs q;
    q.hg = hg_dev;
    q.n = n;
    q.ld = ld;
    q.lambda = (double)st->lambda_diag;
    q.alpha = st->step_length_optim;
    q.max_step = st->max_step;
    q.W = ctx->d_chol.p;
    q.step = step;
    q.step2 = step2;
    q.tail = tail;
    q.flag = reinterpret_cast<int*>(ctx->d_chol.p + (size_t)(n + 1) * ld);
    q.dinv = ctx->d_chol.p + (size_t)(n + 1) * ld + 8;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
    const int tiles = ((n + 31) / 32) * ((n + 31) / 32 + 1) / 2;
    const unsigned grid = (unsigned)std::max(1, std::min(nsm, tiles));  // one trailing-update round per block column where possible
    void* args[] = {&q};
    ProfScope prof_(ctx, PROF_LM_SOLVE);
    CK(cudaLaunchCooperativeKernel((const void*)k_chol_solve, dim3(grid), dim3(CHOL_T), args, 0, ctx->stream));
    ctx->launches++;
    return 0;
}
// which solver a loop body of this context uses: 0 device LU + explicit inverse (P <= 128), 2 device Cholesky, 1 host
inline int bodySolver(const dmsa_b200_ctx* ctx, int P, bool forceHost = false) {
    if (forceHost) return 1;
    if (ctx->solverMode == 2 && P <= CHOL_MAXN) return 2;
    if (ctx->solverMode == 0 && P <= LM_DEV_MAXN) return 0;
    return 1;
}
int solveDev(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, int P, int which) {
    if (which == 2) return cholSolveDev(ctx, st, P, ctx->d_hg.p, ctx->d_step.p, ctx->d_iter.p + 16, ctx->d_iter.p + 16 + P);
    return lmSolveDev(ctx, st, P, ctx->d_hg.p, ctx->d_iter.p + 16, ctx->d_iter.p + 16 + P);
}

// H.diag += lambda; step = -alpha * H^-1 * (J^T e0); NaN guard; infinity-norm clamp      DmsaOptimizer.h:107-128
// returns 1 if the step contains NaN
int solveStep(const dmsa_b200_settings* st, const double* hg, int P, std::vector<double>& step, bool explicit_inverse = true) {
    step.assign(P, 0.0);
    bool nan = false;
    if (explicit_inverse) {  // the reference's expression: (-alpha * H.inverse()) * (J^T e)
        nan = dmsa_host_lm_step(hg, P, (double)st->lambda_diag, st->step_length_optim, step.data());
    } else {
        std::vector<double> H(hg, hg + (size_t)P * P);
        const double* g = hg + (size_t)P * P;
        for (int i = 0; i < P; ++i) H[(size_t)i * P + i] += (double)st->lambda_diag;
        std::vector<double> x;
        if (!chol_solve_vec(H, P, g, x)) lu_solve_vec(H, P, g, x);
        for (int a = 0; a < P; ++a) {
            step[a] = -st->step_length_optim * x[a];
            if (std::isnan(step[a])) nan = true;
        }
    }
    if (nan) return 1;
    double mx = -std::numeric_limits<double>::infinity(), mn = std::numeric_limits<double>::infinity();
    for (double v : step) {
        mx = std::max(mx, v);
        mn = std::min(mn, v);
    }
    const double maxElem = std::max(mx, -mn);
    if (maxElem > st->max_step)
        for (double& v : step) v = (st->max_step / maxElem) * v;
    return 0;
}

int updateGlobalPointsImpl(dmsa_b200_ctx* ctx) {
    if (ctx->model == MODEL_NONE) ARGFAIL("no model staged");
    ctx->poses.relative2global();
    CKRC(uploadParams(ctx));
    const int P = 6 * (ctx->poses.n - 1);
    if (P > 0) CK(cudaMemcpyAsync(ctx->d_batch.p, ctx->d_p.p, P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CKRC(runPoseTables(ctx, 1));
    return transformBase(ctx);
}

// Reference quirk that changes results: ContinuousTrajectory::setPoseParameters (ContinuousTrajectory.h:124-127) only
// writes the RELATIVE poses; the global poses are refreshed by the next updateGlobalPoints().  Hence, when the loop
// ends, decentralize() (:89-93) runs global2relative() on the global poses of the LAST COST EVALUATION (the k = 9
// line-search trial, or the last forward-difference vector on the NaN path) and overwrites the accepted relative
// poses with them.  We keep the host pose state exactly like that: global poses <- chain(p_last_eval), relative
// poses <- p_current.  MapManagement::setPoseParameters (MapManagement.h:197-202) refreshes the chain itself.
void staleGlobal(dmsa_b200_ctx* ctx, const double* p_last_eval, const double* p_current) {
    if (ctx->model == MODEL_KF) {
        ctx->poses.setParams(p_current);
        ctx->poses.relative2global();
        return;
    }
    ctx->poses.setParams(p_last_eval);
    ctx->poses.relative2global();
    ctx->poses.setParams(p_current);
}

// sum over the ranks of a row-sharded set, in place, on the context's stream; no-op without a communicator
int allReduceSum(dmsa_b200_ctx* ctx, double* buf, size_t count) {
    if (!ctx->comm) {
        if (ctx->world > 1) ARGFAIL("the context is sharded (set_shard) but has no communicator: iteration / optimize need dmsa_b200_comm_init; "
                                    "without one, combine the partial results of cost_jacobian_dev / line_search_costs_dev yourself");
        return 0;
    }
    pdlBreak();
    ncclResult_t r = nccl_api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm, ctx->stream);
    if (r != ncclSuccess) {
        ctx->err = "ncclAllReduce: " + nccl_api().describe(r);
        return DMSA_B200_ERR_CUDA;
    }
    ctx->collectives++;
    return 0;
}

// One loop body of optimizeSet (DmsaOptimizer.h:69-144).  With the device LM solver (P <= 128) the whole body runs on the
// device behind ONE read-back: [set counts / octree depths | 9 trial costs | step | e0^T e0 | NaN flag]; the host only picks
// the line-search winner and keeps the pose bookkeeping of the reference (see staleGlobal).  Grid sizes that depend on the
// number of sets come from the previous build of this context and are verified with the read-back (the rare miss redoes the
// body with a synchronous build).  With the host solver (default, and always for P > 128) the body stops twice more: for
// [H | g] before the solve and for the 9 trial costs.
int iterationImpl(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, int32_t* stop, dmsa_b200_report* rep, double* step_out, double* ls_out) {
    PdlScope pdl_(ctx);
    const int P = 6 * (ctx->poses.n - 1);
    if (P <= 0) ARGFAIL("need at least two poses");
    *stop = DMSA_B200_STOP_MAX_ITER;
    struct WallTimer {
        dmsa_b200_ctx* c;
        int id;
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~WallTimer() {
            if (c->profiling) {
                c->profMs[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                c->profCount[id]++;
            }
        }
    } wall_{ctx, PROF_HOST_ITER};
    bool forceHost = false;
    int which = bodySolver(ctx, P);
    bool devSolve = which != 1;  // 0: device LU (P <= 128), 2: device Cholesky (opt-in, any P <= 1024), 1: host solver
    CKRC(ensurePinned(ctx, P));
    CK(ctx->d_hg.ensure((size_t)P * P + P + 1));
    CK(ctx->d_ls.ensure(16));
    CK(ctx->d_iter.ensure(16 + (size_t)P + 2));
    std::vector<double> paramVec, step;
    int nanStep = 0;
    double error0 = 0;
    double ls[9];
    bool tryDefer = devSolve;
    for (int attempt = 0; attempt < 3; ++attempt) {
        which = bodySolver(ctx, P, forceHost);
        devSolve = which != 1;
        // getPoseParameters + updateGlobalPoints at the base pose (the forward-difference batch's vector 0)   :72-75
        ctx->poses.relative2global();
        CKRC(uploadParams(ctx));
        paramVec = ctx->h_p;
        CKRC(prepareFdBatch(ctx));
        CKRC(transformBase(ctx));
        CKRC(buildSets(ctx, st, tryDefer));  // :78-86, :96
        const bool deferred = ctx->G < 0;
        if (!deferred) {
            if (rep) {
                rep->num_gaussians = ctx->G;
                rep->num_extra = numExtra(ctx);
            }
            if (ctx->G < st->min_num_gaussians) {  // :89-93
                *stop = DMSA_B200_STOP_FEW_GAUSSIANS;
                return 0;
            }
        }
        CKRC(runCost(ctx));  // e0 and the P perturbed evaluations in one batch   :99-104
        CKRC(jtjInto(ctx, ctx->d_hg.p));  // :107
        CKRC(allReduceSum(ctx, ctx->d_hg.p, (size_t)P * P + P + 1));
        if (devSolve) {
            // device-resident LM step: solve, line search and ONE read-back
            CKRC(solveDev(ctx, st, P, which));
            CKRC(lineSearchDev(ctx, ctx->d_iter.p));
            CKRC(allReduceSum(ctx, ctx->d_iter.p, 9));
            CK(cudaMemcpyAsync(pinHg(ctx), ctx->d_iter.p, (16 + (size_t)P + 2) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (deferred) CK(cudaMemcpyAsync(pinLinfo(ctx, P), ctx->d_linfo.p, 2 * sizeof(LevelInfo), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (deferred) {
                // late verification of the two guesses the deferred build ran on: octree depth (sort key bits) and set count (grids)
                memcpy(ctx->h_linfo, pinLinfo(ctx, P), 2 * sizeof(LevelInfo));
                int G = 0;
                bool miss = false;
                for (int l = 0; l < 2; ++l) {
                    if (!ctx->levelOn[l]) continue;
                    if (ctx->h_linfo[l].error) ARGFAIL("build_sets: octree deeper than 21 levels (extent / resolution too large)");
                    if (ctx->h_linfo[l].depth > ctx->cachedDepth[l]) miss = true;
                    ctx->cachedDepth[l] = ctx->h_linfo[l].depth;
                    G += ctx->h_linfo[l].G;
                }
                if (G > ctx->cellCap) ARGFAIL("build_sets: Gaussian store capacity exceeded");
                if (G > ctx->Gb) miss = true;
                ctx->Gguess = std::max(G, 1);
                if (miss) {  // redo the body with a synchronous build (the state on the host has not been touched yet)
                    if (ctx->profiling) profCollect(ctx);
                    tryDefer = false;
                    continue;
                }
                ctx->G = G;
                if (rep) {
                    rep->num_gaussians = G;
                    rep->num_extra = numExtra(ctx);
                }
                if (G < st->min_num_gaussians) {  // :89-93 (the evaluations behind it ran on the device but change no state)
                    if (ctx->profiling) profCollect(ctx);
                    *stop = DMSA_B200_STOP_FEW_GAUSSIANS;
                    return 0;
                }
            }
            const double* r = pinHg(ctx);
            if (r[16 + P + 1] == 2.0) {  // Cholesky: the system is not numerically positive definite -> the reference's LU on the host
                if (ctx->profiling) profCollect(ctx);
                forceHost = true;
                tryDefer = false;
                continue;
            }
            std::copy(r, r + 9, ls);
            step.assign(r + 16, r + 16 + P);
            error0 = r[16 + P];
            nanStep = r[16 + P + 1] != 0.0;
            ctx->lastErr0 = error0;
        } else {
            ctx->h_hg.resize((size_t)P * P + P + 1);
            CK(cudaMemcpyAsync(pinHg(ctx), ctx->d_hg.p, ctx->h_hg.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            dmsa_host_solver_arm();  // the solve follows this read-back immediately: helper threads spin up while the host waits
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                dmsa_host_solver_disarm();
                ctx->err = "cudaStreamSynchronize failed before the LM solve";
                return DMSA_B200_ERR_CUDA;
            }
            error0 = pinHg(ctx)[(size_t)P * P + P];
            ctx->lastErr0 = error0;
            {
                WallTimer ws{ctx, PROF_HOST_SOLVE};
                nanStep = solveStep(st, pinHg(ctx), P, step);
            }
            dmsa_host_solver_disarm();
        }
        break;
    }
    if (nanStep) {  // :113-122
        // the last cost evaluation of calcNumericJacobian was p + h e_{P-1}: its global poses stay behind (see staleGlobal)
        if (ctx->profiling) {
            CK(cudaStreamSynchronize(ctx->stream));
            profCollect(ctx);
        }
        std::vector<double> plast = paramVec;
        plast[P - 1] += 1.0 * (double)sqrtf(FLT_EPSILON);
        staleGlobal(ctx, plast.data(), paramVec.data());
        *stop = DMSA_B200_STOP_NAN;
        return 0;
    }
    if (!devSolve) {
        // adaptiveStepSize :152-182 — the 9 trial costs in one batch
        CKRC(lineSearchInto(ctx, step.data(), ctx->d_ls.p));
        CKRC(allReduceSum(ctx, ctx->d_ls.p, 9));
        CK(cudaMemcpyAsync(pinLs(ctx, P), ctx->d_ls.p, 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::copy(pinLs(ctx, P), pinLs(ctx, P) + 9, ls);
    }
    if (ctx->profiling) profCollect(ctx);
    double minError = error0;
    int best = 0;
    for (int k = 1; k < 10; ++k)
        if (ls[k - 1] < minError) {
            minError = ls[k - 1];
            best = k;
        }
    double nrm = 0;
    for (double v : step) nrm += v * v;
    nrm = std::sqrt(nrm);
    if (rep) {
        rep->best_step = best;
        rep->error0 = error0;
        rep->step_norm = nrm;
    }
    if (step_out) std::copy(step.begin(), step.end(), step_out);
    if (ls_out) std::copy(ls, ls + 9, ls_out);
    std::vector<double> pnew(P), plast(P);
    for (int i = 0; i < P; ++i) plast[i] = paramVec[i] + 0.1 * 9.0 * step[i];  // last trial point of adaptiveStepSize
    if (best == 0) {
        // :130-134 — no restore: the set stays at the last trial point p + 0.9*step
        staleGlobal(ctx, plast.data(), plast.data());
        *stop = DMSA_B200_STOP_NO_IMPROVEMENT;
        return 0;
    }
    for (int i = 0; i < P; ++i) pnew[i] = paramVec[i] + 0.1 * (double)best * step[i];
    staleGlobal(ctx, plast.data(), pnew.data());  // :136 setPoseParameters(paramVec)
    if (nrm < st->epsilon) *stop = DMSA_B200_STOP_EPSILON;  // :139-143
    return 0;
}

}  // namespace

// ============================================================================================
extern "C" {

#ifdef DMSA_TIMELINE
int dmsa_b200_dbg_times(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_dbg_t, 16384 * 8); }
#endif
int dmsa_b200_version(void) { return 101; }
int32_t dmsa_b200_fuse_threshold(void) { return FUSE_MAX; }

int dmsa_b200_create(dmsa_b200_ctx** out, int device, void* cuda_stream) {
    if (!out) return DMSA_B200_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return DMSA_B200_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return DMSA_B200_ERR_CUDA;
    dmsa_b200_ctx* ctx = new dmsa_b200_ctx();
    ctx->device = device;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return DMSA_B200_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (ctx->d_flag.ensure(4) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evLevel0, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return DMSA_B200_ERR_CUDA;
    }
    cudaMemsetAsync(ctx->d_flag.p, 0, 4 * sizeof(int), ctx->stream);
    *out = ctx;
    return DMSA_B200_OK;
}

void dmsa_b200_destroy(dmsa_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) {
        nccl_api().CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    if (ctx->stream2) {
        cudaStreamSynchronize(ctx->stream2);
        cudaStreamDestroy(ctx->stream2);
    }
    for (cudaEvent_t e : {ctx->evFork, ctx->evJoin, ctx->evLevel0})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->evScan) cudaEventDestroy(e);
#define REL(b) ctx->b.release()
    REL(d_stamps); REL(d_trajTime); REL(d_urel); REL(d_fh); REL(d_seg); REL(d_hit); REL(d_paramIdx); REL(d_imu); REL(d_kfD); REL(d_plausible);
    REL(d_stage); REL(d_local); REL(d_world); REL(d_normal_l); REL(d_normal_w); REL(d_tid); REL(d_ring); REL(d_flag);
    REL(d_p); REL(d_step); REL(d_batch); REL(d_globO); REL(d_globT); REL(d_quat); REL(d_extra); REL(d_dense); REL(d_Mtab); REL(d_Mpair);
    REL(d_linfo); REL(d_keys); REL(d_bb); REL(d_idx); REL(d_sidx); REL(d_flagA); REL(d_scanA); REL(d_raw_start); REL(d_raw_diff); REL(d_acc_flag);
    REL(d_acc_scan); REL(d_out_cnt); REL(d_sub); REL(d_ntile); REL(d_tile_off); REL(d_best_ij); REL(d_scratch); REL(d_tiles); REL(d_best_v); REL(d_split_nbox); REL(d_split_search); REL(d_code); REL(d_scode); REL(d_ctl); REL(d_rec); REL(d_wrec); REL(d_cell_start); REL(d_cell_n); REL(d_cell_level);
    REL(d_cell_key); REL(d_cell_sub); REL(d_cell_kind); REL(d_okey); REL(d_oval); REL(d_nchunk); REL(d_chunk_off); REL(d_cell_info); REL(d_cell_w0); REL(d_cell_w); REL(d_chunks);
    REL(d_S); REL(d_Q); REL(d_E); REL(d_jpart); REL(d_hg); REL(d_ls); REL(d_lspart); REL(d_done); REL(d_mu); REL(d_biglist); REL(d_gbucket); REL(d_gstart); REL(d_gcursor); REL(d_gcount); REL(d_gpts); REL(d_gquery); REL(d_gsel); REL(d_gctl);
    REL(p_linfo); REL(p_keys); REL(p_bb); REL(p_idx); REL(p_sidx); REL(p_scan); REL(p_raw_start); REL(p_raw_diff); REL(p_pick); REL(p_flag); REL(p_pos); REL(p_rand); REL(p_nn); REL(p_code); REL(p_scode); REL(p_ctl); REL(p_raw); REL(p_out); REL(p_pts); REL(p_cloud); REL(p_range); REL(d_mom); REL(d_solve); REL(d_iter); REL(d_chol);
#undef REL
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->raPin) cudaFreeHost(ctx->raPin);
    for (cudaEvent_t e : ctx->raEv)
        if (e) cudaEventDestroy(e);
    if (ctx->evUpload) cudaEventDestroy(ctx->evUpload);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dmsa_b200_last_error(const dmsa_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t dmsa_b200_launch_count(const dmsa_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int dmsa_b200_synchronize(dmsa_b200_ctx* ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- trajectory model ---------------------------------------------------------------------------
int dmsa_b200_traj_init(dmsa_b200_ctx* ctx, double t_min, double t_max, int32_t n_poses, int32_t use_imu, double dt_res) {
    return dmsa_b200_traj_init_window(ctx, t_min, t_max - t_min + dt_res /* :309 */, n_poses, use_imu, dt_res);
}

// initTraj given the members it leaves behind (t0, horizon): what a binding that sees an already initialised
// ContinuousTrajectory passes, so that `horizon` is the very double the reference computed (no re-derived t_max).
int dmsa_b200_traj_init_window(dmsa_b200_ctx* ctx, double t0, double horizon, int32_t n_poses, int32_t use_imu, double dt_res) {
    if (n_poses < 3) ARGFAIL("traj_init: barycentric_rational of order 2 needs at least 3 control poses");
    if (!(horizon > 0.0) || !(dt_res > 0.0)) ARGFAIL("traj_init: horizon and dt_res must be positive");
    CK(cudaSetDevice(ctx->device));
    ctx->model = MODEL_TRAJ;
    ctx->dt_res = dt_res;
    ctx->useImu = use_imu != 0;
    ctx->imuSet = false;
    ctx->t0 = t0;
    ctx->horizon = horizon;
    ctx->n_total = (int)std::round(ctx->horizon / dt_res) + 1;          // :310
    const int nt = ctx->n_total, n = n_poses;
    // every table below depends on (horizon, n_poses, dt_res) only - all times are relative to t_min - so a window with the
    // same shape as the previous one (the steady state of a sliding window) keeps the tables already in HBM
    // (DMSA_B200_TABLE_CACHE=0 turns the reuse off: bench.py measures its end-to-end leg that way, because it feeds the same
    // window every step and would otherwise never pay for the tables)
    static const bool cacheOn = !(getenv("DMSA_B200_TABLE_CACHE") && atoi(getenv("DMSA_B200_TABLE_CACHE")) == 0);
    if (cacheOn && ctx->tabValid && ctx->tabHorizon == ctx->horizon && ctx->tabPoses == n && ctx->tabDt == dt_res) {
        ctx->poses.resize(n);
        ctx->gravity[0] = 0.0;
        ctx->gravity[1] = 0.0;
        ctx->gravity[2] = -9.805;  // :345
        ctx->n_scan = ctx->n_static = 0;
        return 0;
    }
    ctx->tabValid = false;
    auto linspaced = [](int m, double lo, double hi, std::vector<double>& v) {  // Eigen LinSpaced: lo + i*step, last == hi
        v.resize(m);
        double step = (m > 1) ? (hi - lo) / (double)(m - 1) : 0.0;
        for (int i = 0; i < m; ++i) v[i] = lo + (double)i * step;
        v[m - 1] = hi;
    };
    linspaced(nt, 0.0, ctx->horizon, ctx->trajTime);  // :323
    linspaced(n, 0.0, ctx->horizon, ctx->stamps);     // :332
    ctx->paramIndices.resize(n);
    for (int i = 0; i < n; ++i) ctx->paramIndices[i] = (int)std::round(ctx->stamps[i] / dt_res);  // :336
    ctx->poses.resize(n);
    ctx->gravity[0] = 0.0;
    ctx->gravity[1] = 0.0;
    ctx->gravity[2] = -9.805;  // :345
    ctx->n_scan = ctx->n_static = 0;
    // per-sample interpolation constants: getInterpRotation's index/fraction (:573-581) and the barycentric-rational
    // quotients w_i / (t - s_i) of order 2 (:214; Boost calculate_weights / operator())
    const double* s = ctx->stamps.data();
    std::vector<int> seg(nt), hit(nt);
    std::vector<double> urel(nt), fh((size_t)nt * n), w(n, 0.0);
    const int d = 2;
    for (int k = 0; k < n; ++k) {
        int i_min = std::max(k - d, 0), i_max = (k >= n - d) ? n - d - 1 : k;
        for (int i = i_min; i <= i_max; ++i) {
            double inv_product = 1;
            int j_max = std::min(i + d, n - 1);
            for (int j = i; j <= j_max; ++j) {
                if (j == k) continue;
                inv_product *= (s[k] - s[j]);
            }
            if (i % 2 == 0)
                w[k] += 1 / inv_product;
            else
                w[k] -= 1 / inv_product;
        }
    }
    for (int j = 0; j < nt; ++j) {
        double t = ctx->trajTime[j];
        int r = (int)(std::lower_bound(s, s + n - 1, t) - s);
        seg[j] = r;
        urel[j] = (r != 0) ? (t - s[r - 1]) / (s[r] - s[r - 1]) : 1.0;
        hit[j] = -1;
        for (int i = 0; i < n; ++i) {
            if (t == s[i]) {
                hit[j] = i;
                break;
            }
        }
        for (int i = 0; i < n; ++i) fh[(size_t)j * n + i] = (hit[j] >= 0) ? 0.0 : w[i] / (t - s[i]);
    }
    CK(ctx->d_stamps.ensure(n));
    CK(ctx->d_trajTime.ensure(nt));
    CK(ctx->d_urel.ensure(nt));
    CK(ctx->d_fh.ensure((size_t)nt * n));
    CK(ctx->d_seg.ensure(nt));
    CK(ctx->d_hit.ensure(nt));
    CK(ctx->d_paramIdx.ensure(n));
    CK(cudaMemcpyAsync(ctx->d_stamps.p, s, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_trajTime.p, ctx->trajTime.data(), nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_urel.p, urel.data(), nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_fh.p, fh.data(), (size_t)nt * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_seg.p, seg.data(), nt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_hit.p, hit.data(), nt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_paramIdx.p, ctx->paramIndices.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // host vectors above go out of scope
    ctx->tabValid = true;
    ctx->tabHorizon = ctx->horizon;
    ctx->tabPoses = n;
    ctx->tabDt = dt_res;
    return 0;
}

int dmsa_b200_traj_register_scans(dmsa_b200_ctx* ctx, int32_t n_scans, const dmsa_b200_point_stamp_id* const* scans, const int64_t* sizes,
                                  const float* grid_sizes) {
    if (ctx->model != MODEL_TRAJ) ARGFAIL("register_scans: call traj_init first");
    CK(cudaSetDevice(ctx->device));
    int64_t total = 0;
    float mg = std::numeric_limits<float>::max();
    for (int s = 0; s < n_scans; ++s) {
        total += sizes[s];
        if (grid_sizes) mg = std::min(mg, grid_sizes[s]);  // :235-238
    }
    if (grid_sizes && n_scans > 0) ctx->minGridSize = mg;
    if (total > 0x3fffffff) ARGFAIL("register_scans: too many points");
    const size_t cap = (size_t)total + (size_t)ctx->n_static + 16;
    // static points (if any) are re-appended by the caller after registering (the reference resizes globalPoints, :232)
    ctx->n_static = 0;
    CK(ctx->d_local.ensure(cap));
    CK(ctx->d_world.ensure(cap));
    CK(ctx->d_tid.ensure(cap));
    CK(ctx->d_ring.ensure(cap));
    CK(ctx->d_stage.ensure((size_t)total * 32 + 64));
    CK(cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), ctx->stream));
    // the copies go back to back on stream2 (copy engine), the unpack kernel of scan s waits for its copy only: the
    // unpacking of one scan overlaps the transfer of the next
    while ((int)ctx->evScan.size() < n_scans) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->evScan.push_back(e);
    }
    CK(cudaEventRecord(ctx->evFork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
    int64_t off = 0;
    for (int s = 0; s < n_scans; ++s) {
        if (sizes[s] == 0) continue;
        unsigned char* dst = ctx->d_stage.p + (size_t)off * 32;
        CK(cudaMemcpyAsync(dst, scans[s], (size_t)sizes[s] * 32, cudaMemcpyHostToDevice, ctx->stream2));
        CK(cudaEventRecord(ctx->evScan[s], ctx->stream2));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->evScan[s], 0));
        LAUNCH(k_unpack_psi, cdiv(sizes[s], 256), 256, 0, dst, (int)sizes[s], (int)off, ctx->d_trajTime.p, ctx->n_total, ctx->t0, 0, ctx->d_local.p,
               ctx->d_world.p, ctx->d_ring.p, ctx->d_tid.p, ctx->d_flag.p);
        off += sizes[s];
    }
    ctx->n_scan = total;
    ctx->worldValid = false;
    ctx->worldEpoch++;
    ctx->G = 0;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (flag) {
        ctx->err = "register_scans: a point has homogeneous coordinate w != 1 (PCL_ADD_POINT4D sets 1)";
        return DMSA_B200_ERR_UNSUPPORTED;
    }
    return 0;
}

int dmsa_b200_traj_add_static_points(dmsa_b200_ctx* ctx, const dmsa_b200_point_stamp_id* pts, int64_t n) {
    if (ctx->model != MODEL_TRAJ) ARGFAIL("add_static_points: call traj_init first");
    CK(cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    const size_t newN = (size_t)ctx->n_scan + ctx->n_static + n;
    if (newN > ctx->d_local.cap) {
        // grow, preserving contents
        DBuf<float4> nl, nw;
        DBuf<int> nt, nr;
        CK(nl.ensure(newN));
        CK(nw.ensure(newN));
        CK(nt.ensure(newN));
        CK(nr.ensure(newN));
        const size_t old = (size_t)ctx->n_scan + ctx->n_static;
        if (old) {
            CK(cudaMemcpyAsync(nl.p, ctx->d_local.p, old * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(nw.p, ctx->d_world.p, old * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(nt.p, ctx->d_tid.p, old * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(nr.p, ctx->d_ring.p, old * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->d_local.release();
        ctx->d_world.release();
        ctx->d_tid.release();
        ctx->d_ring.release();
        ctx->d_local = nl;
        ctx->d_world = nw;
        ctx->d_tid = nt;
        ctx->d_ring = nr;
    }
    CK(ctx->d_stage.ensure((size_t)n * 32 + 64));
    CK(cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_stage.p, pts, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_unpack_psi, cdiv(n, 256), 256, 0, ctx->d_stage.p, (int)n, (int)(ctx->n_scan + ctx->n_static), ctx->d_trajTime.p, ctx->n_total, ctx->t0, 1,
           ctx->d_local.p, ctx->d_world.p, ctx->d_ring.p, ctx->d_tid.p, ctx->d_flag.p);
    ctx->n_static += n;
    ctx->worldValid = false;
    ctx->worldEpoch++;
    ctx->G = 0;
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (flag) {
        ctx->err = "add_static_points: a point has homogeneous coordinate w != 1";
        return DMSA_B200_ERR_UNSUPPORTED;
    }
    return 0;
}

int dmsa_b200_traj_remove_static_points(dmsa_b200_ctx* ctx) {
    ctx->n_static = 0;  // :174-187
    ctx->G = 0;         // sets built with the static points are stale
    return 0;
}

int dmsa_b200_traj_get_timing(dmsa_b200_ctx* ctx, int32_t* n_total, double* horizon, double* ctrl_stamps, double* traj_time, int32_t* param_indices) {
    if (ctx->model != MODEL_TRAJ) ARGFAIL("get_timing: no trajectory model");
    if (n_total) *n_total = ctx->n_total;
    if (horizon) *horizon = ctx->horizon;
    if (ctrl_stamps) std::copy(ctx->stamps.begin(), ctx->stamps.end(), ctrl_stamps);
    if (traj_time) std::copy(ctx->trajTime.begin(), ctx->trajTime.end(), traj_time);
    if (param_indices) std::copy(ctx->paramIndices.begin(), ctx->paramIndices.end(), param_indices);
    return 0;
}

int dmsa_b200_traj_get_tform_ids(dmsa_b200_ctx* ctx, int32_t* out) {
    if (ctx->model != MODEL_TRAJ) ARGFAIL("get_tform_ids: no trajectory model");
    CK(cudaMemcpyAsync(out, ctx->d_tid.p, (size_t)ctx->n_scan * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int dmsa_b200_traj_set_imu_factors(dmsa_b200_ctx* ctx, const double* preint_rot, const double* preint_pos, const double* preint_vel,
                                   const double* cov_inv, double balancing_imu, const double* gravity3) {
    if (ctx->model != MODEL_TRAJ) ARGFAIL("set_imu_factors: no trajectory model");
    if (!preint_rot || !preint_pos || !preint_vel || !cov_inv) ARGFAIL("set_imu_factors: null factor array (preint_rot / preint_pos / preint_vel / cov_inv)");
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->poses.n;
    CK(ctx->d_imu.ensure((size_t)n * (9 + 3 + 3 + 81)));
    double* d = ctx->d_imu.p;
    CK(cudaMemcpyAsync(d, preint_rot, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + 9 * (size_t)n, preint_pos, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + 12 * (size_t)n, preint_vel, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + 15 * (size_t)n, cov_inv, (size_t)n * 81 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->balancingImu = balancing_imu;
    if (gravity3)
        for (int a = 0; a < 3; ++a) ctx->gravity[a] = gravity3[a];
    ctx->imuSet = true;
    return 0;
}

// ---- keyframe model -----------------------------------------------------------------------------
int dmsa_b200_kf_init(dmsa_b200_ctx* ctx, int32_t n_keyframes) {
    if (n_keyframes < 2) ARGFAIL("kf_init: need at least two keyframes");
    ctx->model = MODEL_KF;
    ctx->tabValid = false;
    ctx->poses.resize(n_keyframes);
    ctx->kfClouds.assign(n_keyframes, {});
    ctx->kfRings.assign(n_keyframes, {});
    ctx->kfGrid.assign(n_keyframes, 0.3f);
    ctx->useGrav = ctx->useOdom = false;
    ctx->n_scan = ctx->n_static = 0;
    ctx->n_total = 0;
    ctx->gravity[0] = ctx->gravity[1] = 0.0;
    ctx->gravity[2] = -9.805;
    return 0;
}

int dmsa_b200_kf_set_keyframe(dmsa_b200_ctx* ctx, int32_t k, const dmsa_b200_point_normal* pts, const int32_t* ring_ids, int64_t n, float grid_size) {
    if (ctx->model != MODEL_KF || k < 0 || k >= ctx->poses.n) ARGFAIL("kf_set_keyframe: bad keyframe index");
    ctx->kfClouds[k].assign(pts, pts + n);
    ctx->kfRings[k].assign(ring_ids, ring_ids + n);
    ctx->kfGrid[k] = grid_size;
    return 0;
}

int dmsa_b200_kf_commit(dmsa_b200_ctx* ctx) {
    if (ctx->model != MODEL_KF) ARGFAIL("kf_commit: no keyframe model");
    CK(cudaSetDevice(ctx->device));
    int64_t total = 0;
    float mg = std::numeric_limits<float>::max();
    size_t mx = 0;
    for (int k = 0; k < ctx->poses.n; ++k) {
        total += (int64_t)ctx->kfClouds[k].size();
        mx = std::max(mx, ctx->kfClouds[k].size());
        mg = std::min(mg, ctx->kfGrid[k]);  // MapManagement.h:126-131
    }
    if (total > 0x3fffffff) ARGFAIL("kf_commit: too many points");
    ctx->minGridSize = mg;
    CK(ctx->d_local.ensure(total + 16));
    CK(ctx->d_world.ensure(total + 16));
    CK(ctx->d_normal_l.ensure(total + 16));
    CK(ctx->d_normal_w.ensure(total + 16));
    CK(ctx->d_tid.ensure(total + 16));
    CK(ctx->d_ring.ensure(total + 16));
    CK(ctx->d_stage.ensure(mx * 48 + mx * 4 + 64));
    CK(cudaMemsetAsync(ctx->d_flag.p, 0, sizeof(int), ctx->stream));
    int64_t off = 0;
    for (int k = 0; k < ctx->poses.n; ++k) {
        const size_t n = ctx->kfClouds[k].size();
        if (n == 0) continue;
        int* drings = reinterpret_cast<int*>(ctx->d_stage.p + mx * 48);
        CK(cudaMemcpyAsync(ctx->d_stage.p, ctx->kfClouds[k].data(), n * 48, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(drings, ctx->kfRings[k].data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_unpack_pn, cdiv(n, 256), 256, 0, ctx->d_stage.p, drings, (int)n, (int)off, k, ctx->d_local.p, ctx->d_normal_l.p, ctx->d_ring.p, ctx->d_tid.p,
               ctx->d_flag.p);
        CK(cudaStreamSynchronize(ctx->stream));  // staging buffer is reused
        off += (int64_t)n;
    }
    ctx->n_scan = total;
    ctx->n_static = 0;
    ctx->worldValid = false;
    ctx->worldEpoch++;
    ctx->G = 0;
    int flag = 0;
    CK(cudaMemcpy(&flag, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaGetLastError());
    if (flag) {
        ctx->err = "kf_commit: a point has homogeneous coordinate w != 1";
        return DMSA_B200_ERR_UNSUPPORTED;
    }
    return 0;
}

int dmsa_b200_kf_set_gravity_terms(dmsa_b200_ctx* ctx, const double* measured_gravity, const int32_t* plausible, double balance) {
    if (ctx->model != MODEL_KF) ARGFAIL("kf_set_gravity_terms: no keyframe model");
    if (!measured_gravity || !plausible) ARGFAIL("kf_set_gravity_terms: null array");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = ctx->poses.n;
    CK(ctx->d_kfD.ensure((size_t)n * 15));
    CK(ctx->d_plausible.ensure(n));
    CK(cudaMemcpy(ctx->d_kfD.p, measured_gravity, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_plausible.p, plausible, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    ctx->balanceGrav = balance;
    ctx->useGrav = true;
    return 0;
}

int dmsa_b200_kf_set_odometry_terms(dmsa_b200_ctx* ctx, const double* rel_transl, const double* rel_orient_mat, double balance) {
    if (ctx->model != MODEL_KF) ARGFAIL("kf_set_odometry_terms: no keyframe model");
    if (!rel_transl || !rel_orient_mat) ARGFAIL("kf_set_odometry_terms: null array");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = ctx->poses.n;
    CK(ctx->d_kfD.ensure((size_t)n * 15));
    CK(ctx->d_plausible.ensure(n));
    CK(cudaMemcpy(ctx->d_kfD.p + 3 * (size_t)n, rel_transl, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_kfD.p + 6 * (size_t)n, rel_orient_mat, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice));
    ctx->balanceOdom = balance;
    ctx->useOdom = true;
    return 0;
}

// ---- poses ----------------------------------------------------------------------------------------
int dmsa_b200_set_relative_poses(dmsa_b200_ctx* ctx, const double* rel_orient, const double* rel_transl) {
    if (ctx->model == MODEL_NONE) ARGFAIL("set_relative_poses: no model");
    const int n = ctx->poses.n;
    std::copy(rel_orient, rel_orient + 3 * n, ctx->poses.relO.begin());
    std::copy(rel_transl, rel_transl + 3 * n, ctx->poses.relT.begin());
    ctx->poses.relative2global();
    return 0;
}
int dmsa_b200_get_poses(dmsa_b200_ctx* ctx, double* rel_orient, double* rel_transl, double* glob_orient, double* glob_transl) {
    const int n3 = 3 * ctx->poses.n;
    if (rel_orient) std::copy(ctx->poses.relO.begin(), ctx->poses.relO.begin() + n3, rel_orient);
    if (rel_transl) std::copy(ctx->poses.relT.begin(), ctx->poses.relT.begin() + n3, rel_transl);
    if (glob_orient) std::copy(ctx->poses.globO.begin(), ctx->poses.globO.begin() + n3, glob_orient);
    if (glob_transl) std::copy(ctx->poses.globT.begin(), ctx->poses.globT.begin() + n3, glob_transl);
    return 0;
}
int32_t dmsa_b200_num_params(const dmsa_b200_ctx* ctx) { return 6 * (ctx->poses.n - 1); }
int dmsa_b200_get_pose_parameters(dmsa_b200_ctx* ctx, double* params) {
    std::vector<double> p;
    ctx->poses.getParams(p);
    std::copy(p.begin(), p.end(), params);
    return 0;
}
int dmsa_b200_set_pose_parameters(dmsa_b200_ctx* ctx, const double* params) {
    ctx->poses.setParams(params);
    if (ctx->model == MODEL_KF) ctx->poses.relative2global();  // MapManagement.h:197-202
    return 0;
}
int dmsa_b200_centralize(dmsa_b200_ctx* ctx) {
    if (ctx->model != MODEL_TRAJ) return 0;  // MapManagement.h:73-79: early return
    CK(cudaSetDevice(ctx->device));
    for (int a = 0; a < 3; ++a) {
        ctx->origin[a] = ctx->poses.relT[a];
        ctx->poses.relT[a] = 0.0;
    }
    ctx->poses.relative2global();
    if (ctx->n_static > 0) {
        LAUNCH(k_shift_static, cdiv(ctx->n_static, 256), 256, 0, ctx->d_local.p, ctx->d_world.p, (int)ctx->n_scan, (int)ctx->n_static, (float)ctx->origin[0],
               (float)ctx->origin[1], (float)ctx->origin[2], -1);
        CK(cudaGetLastError());
    }
    return 0;
}
int dmsa_b200_decentralize(dmsa_b200_ctx* ctx) {
    if (ctx->model != MODEL_TRAJ) return 0;
    CK(cudaSetDevice(ctx->device));
    ctx->poses.global2relative();
    for (int a = 0; a < 3; ++a) ctx->poses.relT[a] = ctx->origin[a];
    ctx->poses.relative2global();
    if (ctx->n_static > 0) {
        LAUNCH(k_shift_static, cdiv(ctx->n_static, 256), 256, 0, ctx->d_local.p, ctx->d_world.p, (int)ctx->n_scan, (int)ctx->n_static, (float)ctx->origin[0],
               (float)ctx->origin[1], (float)ctx->origin[2], +1);
        CK(cudaGetLastError());
    }
    return 0;
}

// ---- hot path, step by step -------------------------------------------------------------------------
int dmsa_b200_update_global_points(dmsa_b200_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    return updateGlobalPointsImpl(ctx);
}
int64_t dmsa_b200_num_points(const dmsa_b200_ctx* ctx) { return numPoints(ctx); }
int dmsa_b200_get_global_points(dmsa_b200_ctx* ctx, float* xyzw, float* normals) {
    const int64_t N = numPoints(ctx);
    if (xyzw) CK(cudaMemcpyAsync(xyzw, ctx->d_world.p, (size_t)N * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (normals && ctx->model == MODEL_KF) CK(cudaMemcpyAsync(normals, ctx->d_normal_w.p, (size_t)N * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int dmsa_b200_traj_get_dense_tforms(dmsa_b200_ctx* ctx, float* out) {
    if (ctx->model != MODEL_TRAJ || ctx->curVld == 0) ARGFAIL("get_dense_tforms: call update_global_points first");
    // column v = 0 of the table
    CK(cudaMemcpy2DAsync(out, 48, ctx->d_Mtab.p, (size_t)ctx->curVld * 48, 48, ctx->n_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// validation surface: the transform table of the batch evaluated last, [rows + 1][V][12] floats (the last row is the identity
// used by static points); dims = {rows + 1, V}.  out may be NULL (dimensions only).
int dmsa_b200_get_batch_tables(dmsa_b200_ctx* ctx, float* out, int32_t* dims) {
    if (ctx->model == MODEL_NONE || ctx->curVld == 0) ARGFAIL("get_batch_tables: no batch evaluated yet");
    CK(cudaSetDevice(ctx->device));
    const int rows1 = numTableRows(ctx) + 1, V = ctx->curV, Vld = ctx->curVld;
    if (dims) {
        dims[0] = rows1;
        dims[1] = V;
    }
    if (!out) return 0;
    CK(cudaMemcpy2DAsync(out, (size_t)V * 48, ctx->d_Mtab.p, (size_t)Vld * 48, (size_t)V * 48, rows1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// denseGlobalPoses of the current parameters (ContinuousTrajectory.h:194-218): 3 x n_total column-major doubles each
int dmsa_b200_traj_get_dense_poses(dmsa_b200_ctx* ctx, double* orient, double* transl) {
    if (ctx->model != MODEL_TRAJ || ctx->curVld == 0) ARGFAIL("get_dense_poses: call update_global_points first");
    CK(cudaSetDevice(ctx->device));
    const int nt = ctx->n_total;
    CK(ctx->d_dense.ensure((size_t)6 * nt));
    PoseBatch pb;
    memset(&pb, 0, sizeof(pb));
    pb.V = ctx->curV;
    pb.Vld = ctx->curVld;
    pb.n = ctx->poses.n;
    pb.globO_t = ctx->d_globO.p;
    pb.globT_t = ctx->d_globT.p;
    pb.quat_t = ctx->d_quat.p;
    TrajTiming tt;
    tt.n_total = nt;
    tt.seg = ctx->d_seg.p;
    tt.urel = ctx->d_urel.p;
    tt.fh = ctx->d_fh.p;
    tt.hit = ctx->d_hit.p;
    LAUNCH(k_dense_poses, cdiv(nt, 128), 128, 0, pb, tt, 0, ctx->d_dense.p, ctx->d_dense.p + (size_t)3 * nt);
    if (orient) CK(cudaMemcpyAsync(orient, ctx->d_dense.p, (size_t)3 * nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (transl) CK(cudaMemcpyAsync(transl, ctx->d_dense.p + (size_t)3 * nt, (size_t)3 * nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

int dmsa_b200_build_sets(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, int32_t* num_gaussians, int64_t* num_memberships) {
    PdlScope pdl_(ctx);
    CK(cudaSetDevice(ctx->device));
    CKRC(buildSets(ctx, settings));
    if (num_gaussians) *num_gaussians = ctx->G;
    if (num_memberships) {
        std::vector<int> n(ctx->G);
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->G) CK(cudaMemcpy(n.data(), ctx->d_cell_n.p, (size_t)ctx->G * sizeof(int), cudaMemcpyDeviceToHost));
        int64_t M = 0;
        for (int v : n) M += v;
        ctx->M = M;
        *num_memberships = M;
    }
    return 0;
}

int dmsa_b200_get_sets(dmsa_b200_ctx* ctx, int64_t* offsets, int32_t* members, float* info, float* weights, int32_t* level, int32_t* key, int32_t* sub) {
    const int G = ctx->G;
    const int64_t N = numPoints(ctx);
    std::vector<int> start(G), n(G);
    // the set statistics run asynchronously on the context's (non-blocking) streams; stream2 is joined into ctx->stream by
    // buildSets, so one synchronisation here orders every blocking copy below behind them
    CK(cudaStreamSynchronize(ctx->stream));
    if (G) {
        CK(cudaMemcpy(start.data(), ctx->d_cell_start.p, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(n.data(), ctx->d_cell_n.p, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (offsets) {
        offsets[0] = 0;
        for (int g = 0; g < G; ++g) offsets[g + 1] = offsets[g] + n[g];
    }
    if (members && G) {
        std::vector<int> sidx((size_t)2 * N);
        CK(cudaMemcpy(sidx.data(), ctx->d_sidx.p, (size_t)2 * N * sizeof(int), cudaMemcpyDeviceToHost));
        int64_t o = 0;
        for (int g = 0; g < G; ++g) {
            std::copy(sidx.begin() + start[g], sidx.begin() + start[g] + n[g], members + o);
            o += n[g];
        }
    }
    if (info && G) CK(cudaMemcpy(info, ctx->d_cell_info.p, (size_t)G * 9 * sizeof(float), cudaMemcpyDeviceToHost));
    if (weights && G) CK(cudaMemcpy(weights, ctx->d_cell_w.p, (size_t)G * sizeof(float), cudaMemcpyDeviceToHost));
    if (level && G) CK(cudaMemcpy(level, ctx->d_cell_level.p, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost));
    if (key && G) CK(cudaMemcpy(key, ctx->d_cell_key.p, (size_t)G * 3 * sizeof(int), cudaMemcpyDeviceToHost));
    if (sub && G) CK(cudaMemcpy(sub, ctx->d_cell_sub.p, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int dmsa_b200_get_voxel_keys(dmsa_b200_ctx* ctx, int32_t level, int32_t* keys, int64_t* root_lo, int32_t* depth) {
    if (level < 0 || level > 1 || !ctx->levelOn[level]) ARGFAIL("get_voxel_keys: level not built");
    const int64_t N = numPoints(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    if (keys) CK(cudaMemcpy(keys, ctx->d_keys.p + (size_t)3 * N * level, (size_t)3 * N * sizeof(int), cudaMemcpyDeviceToHost));
    if (root_lo)
        for (int a = 0; a < 3; ++a) root_lo[a] = ctx->h_linfo[level].lo[a];
    if (depth) *depth = ctx->h_linfo[level].depth;
    return 0;
}

int dmsa_b200_eval_cost(dmsa_b200_ctx* ctx, const double* params, int32_t n_vectors, double* e) {
    CK(cudaSetDevice(ctx->device));
    const int P = 6 * (ctx->poses.n - 1), V = n_vectors;
    if (V < 1) ARGFAIL("eval_cost: n_vectors < 1");
    if (ctx->G <= 0) ARGFAIL("eval_cost: build_sets first");
    CK(ctx->d_batch.ensure((size_t)std::max(V, P + 1) * P));
    CK(cudaMemcpyAsync(ctx->d_batch.p, params, (size_t)V * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CKRC(runPoseTables(ctx, V));
    CKRC(runCost(ctx));
    const int R = ctx->G + numExtra(ctx), Vld = ctx->curVld;
    std::vector<double> buf((size_t)R * Vld);
    CK(cudaMemcpyAsync(buf.data(), ctx->d_E.p, buf.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < V; ++v)
        for (int r = 0; r < R; ++r) e[(size_t)v * R + r] = buf[(size_t)r * Vld + v];
    return 0;
}

int dmsa_b200_cost_jacobian(dmsa_b200_ctx* ctx, double* H, double* g, double* err0, double* e0, double* J) {
    PdlScope pdl_(ctx);
    CK(cudaSetDevice(ctx->device));
    const int P = 6 * (ctx->poses.n - 1);
    if (ctx->G <= 0) ARGFAIL("cost_jacobian: build_sets first");
    CKRC(uploadParams(ctx));
    CKRC(prepareFdBatch(ctx));
    CKRC(runCost(ctx));
    CK(ctx->d_hg.ensure((size_t)P * P + P + 1));
    CKRC(jtjInto(ctx, ctx->d_hg.p));
    ctx->h_hg.resize((size_t)P * P + P + 1);
    CK(cudaMemcpyAsync(ctx->h_hg.data(), ctx->d_hg.p, ctx->h_hg.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    const int R = ctx->G + numExtra(ctx), Vld = ctx->curVld;
    std::vector<double> buf;
    if (e0 || J) {
        buf.resize((size_t)R * Vld);
        CK(cudaMemcpyAsync(buf.data(), ctx->d_E.p, buf.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (H) std::copy(ctx->h_hg.begin(), ctx->h_hg.begin() + (size_t)P * P, H);
    if (g) std::copy(ctx->h_hg.begin() + (size_t)P * P, ctx->h_hg.begin() + (size_t)P * P + P, g);
    if (err0) *err0 = ctx->h_hg[(size_t)P * P + P];
    if (e0)
        for (int r = 0; r < R; ++r) e0[r] = buf[(size_t)r * Vld];
    if (J) {
        const double inv_h = 1.0 / (1.0 * (double)sqrtf(FLT_EPSILON));
        for (int k = 0; k < P; ++k)
            for (int r = 0; r < R; ++r) J[(size_t)k * R + r] = inv_h * (buf[(size_t)r * Vld + k + 1] - buf[(size_t)r * Vld]);
    }
    return 0;
}

int dmsa_b200_iteration(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, int32_t* stop, dmsa_b200_report* report, double* step, double* ls_cost) {
    CK(cudaSetDevice(ctx->device));
    int32_t s = 0;
    dmsa_b200_report rep;
    memset(&rep, 0, sizeof(rep));
    int rc = iterationImpl(ctx, settings, &s, &rep, step, ls_cost);
    rep.iterations = 1;
    rep.stop_reason = s;
    if (stop) *stop = s;
    if (report) *report = rep;
    return rc;
}

}  // extern "C"
namespace {
// One loop body enqueued without any host synchronisation (device LM solver, deferred set build, winner / update / stop
// tests by k_iter_decide), its read-back block copied to `slot` of the pinned ring.  upload: d_p <- the host's parameters
// (first run-ahead body); otherwise d_p is what the previous body's k_iter_decide left.
int enqueueBody(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, int P, bool upload, double* slot, cudaEvent_t done) {
    PdlScope pdl_(ctx);
    const size_t recN = 16 + 3 * (size_t)P + 2;
    if (upload) CKRC(uploadParams(ctx));
    CKRC(prepareFdBatch(ctx));
    CKRC(transformBase(ctx));
    CKRC(buildSets(ctx, st, true));
    if (ctx->G >= 0) ARGFAIL("optimize: the run-ahead loop needs a deferred set build");
    CKRC(runCost(ctx));
    CKRC(jtjInto(ctx, ctx->d_hg.p));
    CKRC(allReduceSum(ctx, ctx->d_hg.p, (size_t)P * P + P + 1));
    CKRC(solveDev(ctx, st, P, bodySolver(ctx, P)));
    CKRC(lineSearchDev(ctx, ctx->d_iter.p));
    CKRC(allReduceSum(ctx, ctx->d_iter.p, 9));
    LAUNCH(k_iter_decide, 1, 128, 0, ctx->d_iter.p, ctx->d_p.p, P, st->epsilon);
    CK(cudaMemcpyAsync(slot, ctx->d_iter.p, recN * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(slot + recN, ctx->d_linfo.p, 2 * sizeof(LevelInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(done, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}
}  // namespace
extern "C" {

int dmsa_b200_optimize(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, dmsa_b200_report* report) {
    PdlScope pdl_(ctx);
    CK(cudaSetDevice(ctx->device));
    if (ctx->model == MODEL_NONE) ARGFAIL("optimize: no model staged");
    dmsa_b200_report rep;
    memset(&rep, 0, sizeof(rep));
    if (settings->use_centralization) CKRC(dmsa_b200_centralize(ctx));  // :66-67
    int it = 0;
    int32_t stop = DMSA_B200_STOP_MAX_ITER;
    const int P = 6 * (ctx->poses.n - 1);
    // Run-ahead loop (device LM solver only): body i + 1 is enqueued before the host has read body i's results, so the GPU
    // never waits for the host between bodies.  The host consumes the read-back blocks one body late and keeps the
    // reference's pose bookkeeping (staleGlobal); a body that ran past a stop condition is discarded.
    const bool runAhead = ctx->runAhead && P > 0 && bodySolver(ctx, P) != 1 && !ctx->profiling && settings->num_iter >= 3;
    if (runAhead) {
        // body 0 takes the synchronous path (its set build has no size guess yet)
        CKRC(iterationImpl(ctx, settings, &stop, &rep, nullptr, nullptr));
        it = 1;
        const size_t recN = 16 + 3 * (size_t)P + 2;
        const size_t slotBytes = recN * sizeof(double) + 2 * sizeof(LevelInfo);
        if (stop == DMSA_B200_STOP_MAX_ITER && it < settings->num_iter) {
            if (ctx->raCap < 2 * slotBytes) {
                if (ctx->raPin) cudaFreeHost(ctx->raPin);
                ctx->raPin = nullptr;
                CK(cudaHostAlloc((void**)&ctx->raPin, 2 * slotBytes, cudaHostAllocDefault));
                ctx->raCap = 2 * slotBytes;
            }
            for (int k = 0; k < 2; ++k)
                if (!ctx->raEv[k]) CK(cudaEventCreateWithFlags(&ctx->raEv[k], cudaEventDisableTiming));
            CK(ctx->d_hg.ensure((size_t)P * P + P + 1));
            CK(ctx->d_iter.ensure(recN + 8));
            auto slotPtr = [&](int b) { return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ctx->raPin) + (size_t)(b & 1) * slotBytes); };
            int enq = it;  // next body to enqueue
            ctx->poses.relative2global();
            CKRC(enqueueBody(ctx, settings, P, true, slotPtr(enq), ctx->raEv[enq & 1]));
            ++enq;
            bool fallback = false;
            while (it < settings->num_iter) {
                if (enq < settings->num_iter && enq == it + 1) {  // keep one body in flight behind the one being consumed
                    CKRC(enqueueBody(ctx, settings, P, false, slotPtr(enq), ctx->raEv[enq & 1]));
                    ++enq;
                }
                CK(cudaEventSynchronize(ctx->raEv[it & 1]));
                const double* r = slotPtr(it);
                LevelInfo li[2];
                memcpy(li, r + recN, 2 * sizeof(LevelInfo));
                // late verification of the guesses the deferred build ran on (octree depth, set count)
                int G = 0;
                bool miss = false;
                for (int l = 0; l < 2; ++l) {
                    if (!ctx->levelOn[l]) continue;
                    if (li[l].error) ARGFAIL("build_sets: octree deeper than 21 levels (extent / resolution too large)");
                    if (li[l].depth > ctx->cachedDepth[l]) miss = true;
                    G += li[l].G;
                }
                if (G > ctx->cellCap) ARGFAIL("build_sets: Gaussian store capacity exceeded");
                if (G > std::max(li[0].bound, li[1].bound)) miss = true;
                if (r[16 + P + 1] == 2.0) miss = true;  // Cholesky refused the system: this body goes to the host LU
                if (miss) {  // redo this body (and the rest) on the synchronous path; the host state has not been touched by it
                    CK(cudaStreamSynchronize(ctx->stream));
                    for (int l = 0; l < 2; ++l)
                        if (ctx->levelOn[l]) ctx->cachedDepth[l] = std::max(ctx->cachedDepth[l], li[l].depth);
                    fallback = true;
                    break;
                }
                for (int l = 0; l < 2; ++l)
                    if (ctx->levelOn[l]) ctx->cachedDepth[l] = li[l].depth;
                ctx->Gguess = std::max(G, 1);
                memcpy(ctx->h_linfo, li, sizeof(li));
                rep.num_gaussians = G;
                rep.num_extra = numExtra(ctx);
                ++it;
                if (G < settings->min_num_gaussians) {  // :89-93
                    stop = DMSA_B200_STOP_FEW_GAUSSIANS;
                    break;
                }
                const double error0 = r[16 + P];
                ctx->lastErr0 = error0;
                const int code = (int)r[10];
                rep.error0 = error0;
                if (code == DMSA_B200_STOP_NAN) {  // :113-122
                    std::vector<double> cur;
                    ctx->poses.getParams(cur);
                    std::vector<double> plast = cur;
                    plast[P - 1] += 1.0 * (double)sqrtf(FLT_EPSILON);
                    staleGlobal(ctx, plast.data(), cur.data());
                    stop = DMSA_B200_STOP_NAN;
                    break;
                }
                rep.best_step = (int)r[9];
                rep.step_norm = r[11];
                staleGlobal(ctx, r + 16 + 2 * P + 2, r + 16 + P + 2);  // (last trial point, accepted parameters)
                if (code != DMSA_B200_STOP_MAX_ITER) {
                    stop = code;
                    break;
                }
            }
            ctx->G = rep.num_gaussians;
            if (fallback) {
                for (; it < settings->num_iter; ++it) {
                    CKRC(iterationImpl(ctx, settings, &stop, &rep, nullptr, nullptr));
                    if (stop != DMSA_B200_STOP_MAX_ITER) {
                        ++it;
                        break;
                    }
                }
            }
        }
    } else {
    for (; it < settings->num_iter; ++it) {  // :69
        CKRC(iterationImpl(ctx, settings, &stop, &rep, nullptr, nullptr));
        if (stop != DMSA_B200_STOP_MAX_ITER) {
            ++it;
            break;
        }
    }
    }
    if (settings->use_centralization) CKRC(dmsa_b200_decentralize(ctx));  // :146-147
    CKRC(updateGlobalPointsImpl(ctx));                                     // :149
    CK(cudaStreamSynchronize(ctx->stream));
    rep.iterations = it;
    rep.stop_reason = stop;
    if (report) *report = rep;
    return 0;
}

int dmsa_b200_set_mean_mode(dmsa_b200_ctx* ctx, int32_t mode) {
    if (mode != 0 && mode != 1) ARGFAIL("set_mean_mode: 0 (order-free, default) or 1 (reference-sequential)");
    ctx->meanMode = mode;
    return 0;
}

// ---- per-kernel timing (CUDA events on the launching stream) ------------------------------------------------
int dmsa_b200_profile_enable(dmsa_b200_ctx* ctx, int32_t on) {
    ctx->profiling = on != 0;
    for (int i = 0; i < PROF_NUM; ++i) {
        ctx->profMs[i] = 0;
        ctx->profCount[i] = 0;
    }
    return 0;
}
int32_t dmsa_b200_profile_num(void) { return PROF_NUM; }
const char* dmsa_b200_profile_name(int32_t id) { return (id >= 0 && id < PROF_NUM) ? kProfNames[id] : ""; }
int dmsa_b200_profile_read(dmsa_b200_ctx* ctx, int32_t id, double* total_ms, int64_t* count) {
    if (id < 0 || id >= PROF_NUM) ARGFAIL("profile_read: bad id");
    CK(cudaStreamSynchronize(ctx->stream));
    profCollect(ctx);
    if (total_ms) *total_ms = ctx->profMs[id];
    if (count) *count = ctx->profCount[id];
    return 0;
}

// Host-side LM step of DmsaOptimizer.h:107-128 on an (all-reduced) [H | g | err0] buffer: H.diag += lambda,
// step = -alpha * H^-1 * g (explicit LU inverse like Eigen's inverse()), NaN guard, infinity-norm clamp.
int dmsa_b200_lm_solve(const dmsa_b200_settings* settings, const double* hg, int32_t n_params, int32_t explicit_inverse, double* step,
                       int32_t* has_nan) {
    if (!settings || !hg || !step || n_params <= 0) return DMSA_B200_ERR_ARG;
    std::vector<double> st;
    if (explicit_inverse == 2) dmsa_host_solver_arm();  // same arithmetic, substitution columns spread over the helper threads
    int nan = solveStep(settings, hg, n_params, st, explicit_inverse != 0);
    if (explicit_inverse == 2) dmsa_host_solver_disarm();
    std::copy(st.begin(), st.end(), step);
    if (has_nan) *has_nan = nan;
    return 0;
}

// Which solver dmsa_b200_iteration / dmsa_b200_optimize use for the LM step: 0 (default) the device kernels (P <= 128: the
// loop body then runs behind one read-back; larger systems use the host solver in either mode), 1 the host solver.
// Same operation sequence, bit-identical steps.
int dmsa_b200_set_lm_solver(dmsa_b200_ctx* ctx, int32_t mode) {
    if (mode < 0 || mode > 2) ARGFAIL("set_lm_solver: 0 (device LU + explicit inverse for P <= 128, default), 1 (host) or 2 (device Cholesky, P <= 1024)");
    ctx->solverMode = mode;
    return 0;
}

// dmsa_b200_optimize: 1 (default) run-ahead loop — body i + 1 is enqueued before the host reads body i's results; 0: one
// body at a time.  Same results bit for bit.
int dmsa_b200_set_run_ahead(dmsa_b200_ctx* ctx, int32_t on) {
    ctx->runAhead = on != 0;
    return 0;
}
// Cost kernels of the forward-difference batch: 1 (default) pair-packed FP32x2 kernels, 0 scalar kernels (bit-identical).
int dmsa_b200_set_pair_mode(dmsa_b200_ctx* ctx, int32_t mode) {
    if (mode < 0 || mode > 2) ARGFAIL("set_pair_mode: 1 (pair-packed with the shared-rotation fast path, default), 2 (pair-packed, every vector in full) or 0 (scalar)");
    ctx->pairMode = mode;
    return 0;
}

// The device LM step on a HOST copy of [H | g | err0] (n_params <= 128): upload, the three solver kernels, download.
// Exists so that the device solver can be checked against dmsa_b200_lm_solve (host) on arbitrary systems.
int dmsa_b200_lm_solve_device(dmsa_b200_ctx* ctx, const dmsa_b200_settings* settings, const double* hg, int32_t n_params, double* step,
                              int32_t* has_nan) {
    if (!settings || !hg || !step || n_params <= 0) ARGFAIL("lm_solve_device: bad arguments");
    if (n_params > LM_DEV_MAXN) ARGFAIL("lm_solve_device: the device solver takes at most 128 parameters (larger systems: host solver)");
    CK(cudaSetDevice(ctx->device));
    const int P = n_params;
    const size_t nhg = (size_t)P * P + P + 1;
    CK(ctx->d_hg.ensure(nhg));
    CK(ctx->d_step.ensure(P));
    CK(ctx->d_iter.ensure(16 + (size_t)P + 2));
    CK(cudaMemcpyAsync(ctx->d_hg.p, hg, nhg * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CKRC(lmSolveDev(ctx, settings, P, ctx->d_hg.p, ctx->d_iter.p + 16, ctx->d_iter.p + 16 + P));
    std::vector<double> r((size_t)P + 2);
    CK(cudaMemcpyAsync(r.data(), ctx->d_iter.p + 16, r.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::copy(r.begin(), r.begin() + P, step);
    if (has_nan) *has_nan = r[(size_t)P + 1] != 0.0;
    return 0;
}

// ---- static-point selection and overlap ratio (DmsaSlam.h:264-414) ---------------------------------
namespace {
// hashed grid over `n` points (device; every stride4-th float4 is a point) with cell edge h (kernels_knn.cuh)
int buildHashGrid(dmsa_b200_ctx* ctx, const float4* pts, int n, int stride4, double h, HashGrid* view) {
    PdlScope pdl_(ctx);
    if (!(h > 0.0)) h = 1e-6;
    int B = 4096;
    while (B < 2 * n && B < (1 << 22)) B <<= 1;
    view->n = n;
    view->B = B;
    view->h = h;
    view->start = nullptr;
    view->pts = nullptr;
    if (n == 0) return 0;
    CK(ctx->d_gbucket.ensure(n));
    CK(ctx->d_gstart.ensure((size_t)B + 2));
    CK(ctx->d_gcursor.ensure((size_t)2 * B + 2));  // [counts (B + 1) | cursors (B)]
    CK(ctx->d_gpts.ensure(n));
    const int tiles = (B + 1 + CS_TILE - 1) / CS_TILE;
    const size_t ctlBytes = 16 + (size_t)tiles * sizeof(u64_t);
    CK(ctx->d_gctl.ensure(ctlBytes));
    CK(cudaMemsetAsync(ctx->d_gctl.p, 0, ctlBytes, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_gcursor.p, 0, ((size_t)2 * B + 2) * sizeof(int), ctx->stream));
    int* counts = ctx->d_gcursor.p;
    int* cursor = ctx->d_gcursor.p + B + 1;
    LAUNCH(k_hg_count, cdiv(n, 256), 256, 0, pts, n, stride4, h, B, ctx->d_gbucket.p, counts);
    ScanArgs sa;
    sa.in = counts;
    sa.out = ctx->d_gstart.p;
    sa.n = B + 1;
    sa.tiles = tiles;
    sa.status = reinterpret_cast<u64_t*>(ctx->d_gctl.p + 16);
    sa.ticket = reinterpret_cast<int*>(ctx->d_gctl.p);
    LAUNCH(k_scan_excl, tiles, CS_T, 0, sa);
    LAUNCH(k_hg_fill, cdiv(n, 256), 256, 0, pts, n, stride4, ctx->d_gbucket.p, ctx->d_gstart.p, cursor, ctx->d_gpts.p);
    view->start = ctx->d_gstart.p;
    view->pts = ctx->d_gpts.p;
    CK(cudaGetLastError());
    return 0;
}
// the radius grid of the staged window cloud (cell edge radius * 1.000001), rebuilt only when the cloud or the radius changed
int windowRadiusGrid(dmsa_b200_ctx* ctx, float radius, HashGrid* view) {
    const int N = (int)numPoints(ctx);
    if (ctx->gridEpoch == ctx->worldEpoch && ctx->gridRadius == radius && ctx->gridN == N && N > 0) {
        view->n = N;
        view->B = ctx->gridB;
        view->h = ctx->gridH;
        view->start = ctx->d_gstart.p;
        view->pts = ctx->d_gpts.p;
        return 0;
    }
    CKRC(buildHashGrid(ctx, ctx->d_world.p, N, 1, (double)radius * 1.000001, view));
    ctx->gridEpoch = ctx->worldEpoch;
    ctx->gridRadius = radius;
    ctx->gridN = N;
    ctx->gridB = view->B;
    ctx->gridH = view->h;
    return 0;
}
}  // namespace

// addStaticPoints, inner loop for ONE keyframe cloud (DmsaSlam.h:304-339): selected[j] = 1 iff the nearest point of the
// staged window cloud (globalPoints as of the last update_global_points / add_static_points) is within max_dist and the
// point is visible from pos; *num_selected = currOverlap.  The window grid is built by the first call and kept for the
// following keyframe clouds as long as the window cloud and the radius stay the same (the reference builds its kd-tree once
// per addStaticPoints call, DmsaSlam.h:283-286).
int dmsa_b200_select_static_points(dmsa_b200_ctx* ctx, const dmsa_b200_point_normal* cloud, int64_t n, const float* pos, float max_dist,
                                   uint8_t* selected, int64_t* num_selected) {
    PdlScope pdl_(ctx);
    if ((n > 0 && (!cloud || !selected)) || !pos || n < 0 || n > 0x3fffffff) ARGFAIL("select_static_points: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (num_selected) *num_selected = 0;
    if (n == 0) return 0;
    const int64_t N = numPoints(ctx);
    if (N > 0 && !ctx->worldValid) ARGFAIL("select_static_points: call update_global_points first (the search runs on globalPoints)");
    HashGrid g;
    CKRC(windowRadiusGrid(ctx, max_dist, &g));
    const float max_sq = (float)std::pow((double)(1.0f * max_dist), 2);  // DmsaSlam.h:293
    CK(ctx->d_gquery.ensure((size_t)3 * n));
    CK(ctx->d_gsel.ensure((size_t)n));
    CK(ctx->d_gcount.ensure(1));
    CK(cudaMemcpyAsync(ctx->d_gquery.p, cloud, (size_t)n * 48, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_gcount.p, 0, sizeof(int), ctx->stream));
    LAUNCH(k_select_static, cdiv(n, 128), 128, 0, g, ctx->d_gquery.p, (int)n, pos[0], pos[1], pos[2], max_sq, ctx->d_gsel.p, ctx->d_gcount.p);
    int cnt = 0;
    CK(cudaMemcpyAsync(selected, ctx->d_gsel.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&cnt, ctx->d_gcount.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (num_selected) *num_selected = cnt;
    return 0;
}

// getOverlap(pc1, pc2 = the staged window cloud, max_dist) (DmsaSlam.h:377-414): pc1 = n1 x (x, y, z, w) floats on the host
int dmsa_b200_overlap(dmsa_b200_ctx* ctx, const float* pc1_xyzw, int64_t n1, float max_dist, float* overlap) {
    PdlScope pdl_(ctx);
    if (!overlap || n1 < 0 || n1 > 0x3fffffff || (n1 > 0 && !pc1_xyzw)) ARGFAIL("overlap: bad arguments");
    CK(cudaSetDevice(ctx->device));
    *overlap = 0.0f;
    const int64_t N = numPoints(ctx);
    if (n1 == 0 || N == 0) return 0;  // :380-381
    if (!ctx->worldValid) ARGFAIL("overlap: call update_global_points first (the search runs on globalPoints)");
    CK(ctx->d_gquery.ensure((size_t)n1));
    CK(cudaMemcpyAsync(ctx->d_gquery.p, pc1_xyzw, (size_t)n1 * 16, cudaMemcpyHostToDevice, ctx->stream));
    HashGrid g;
    ctx->gridEpoch = ~0ull;  // the grid buffers now hold pc1, not the window
    CKRC(buildHashGrid(ctx, ctx->d_gquery.p, (int)n1, 1, (double)max_dist * 1.000001, &g));
    const float max_sq = max_dist * max_dist;  // :386
    CK(ctx->d_gcount.ensure(1));
    CK(cudaMemsetAsync(ctx->d_gcount.p, 0, sizeof(int), ctx->stream));
    LAUNCH(k_overlap_count, cdiv(N, 128), 128, 0, g, ctx->d_world.p, (int)N, max_sq, ctx->d_gcount.p);
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->d_gcount.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    *overlap = static_cast<float>(cnt) / static_cast<float>((int)N);  // :412
    return 0;
}

// ---- multi-GPU row sharding -----------------------------------------------------------------------------
int dmsa_b200_set_shard(dmsa_b200_ctx* ctx, int32_t rank, int32_t world) {
    if (world < 1 || rank < 0 || rank >= world) ARGFAIL("set_shard: bad rank/world");
    if (rank != ctx->rank || world != ctx->world) ctx->G = 0;  // the ownership plan of the last build_sets is stale
    ctx->rank = rank;
    ctx->world = world;
    return 0;
}
// ---- SPD solve of the keyframe-bundle extension (kernels_chol.cuh): device in, device out ----
int dmsa_b200_spd_solve_dev(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, const double* hg_dev, int32_t n, double* step_dev, double* tail_dev) {
    if (!st || !hg_dev || !step_dev || !tail_dev || n <= 0 || n > CHOL_MAXN) ARGFAIL("spd_solve_dev: bad arguments (1 <= n <= 1024)");
    CK(cudaSetDevice(ctx->device));
    return cholSolveDev(ctx, st, n, hg_dev, step_dev, nullptr, tail_dev);
}
// the same on host buffers (validation): step[n], *flag = 0 ok / 1 NaN / 2 not positive definite
int dmsa_b200_spd_solve(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, const double* hg, int32_t n, double* step, int32_t* flag) {
    if (!st || !hg || !step || n <= 0 || n > CHOL_MAXN) ARGFAIL("spd_solve: bad arguments (1 <= n <= 1024)");
    CK(cudaSetDevice(ctx->device));
    const size_t nhg = (size_t)n * n + n + 1;
    CK(ctx->d_hg.ensure(nhg));
    CK(ctx->d_iter.ensure(16 + (size_t)n + 2));
    CK(cudaMemcpyAsync(ctx->d_hg.p, hg, nhg * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CKRC(dmsa_b200_spd_solve_dev(ctx, st, ctx->d_hg.p, n, ctx->d_iter.p + 16, ctx->d_iter.p + 16 + n));
    std::vector<double> r((size_t)n + 2);
    CK(cudaMemcpyAsync(r.data(), ctx->d_iter.p + 16, r.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::copy(r.begin(), r.begin() + n, step);
    if (flag) *flag = (int)r[(size_t)n + 1];
    return 0;
}

// ---- keyframe-bundle extension (BASELINE config 4): one context per bundle, the orchestration of an iteration is
//      jacobian (every bundle) -> all_reduce -> spd_solve_dev -> line_search (every bundle) -> all_reduce -> ONE read-back
//      -> verify (every bundle).  Nothing in between touches the host.
int dmsa_b200_bundle_jacobian(dmsa_b200_ctx* ctx, const dmsa_b200_settings* st, const int32_t* idx_dev, int32_t P_global, double* ghg_dev, int32_t sync_build) {
    PdlScope pdl_(ctx);
    if (!st || !idx_dev || !ghg_dev) ARGFAIL("bundle_jacobian: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (ctx->model == MODEL_NONE) ARGFAIL("bundle_jacobian: no model staged");
    const int P = 6 * (ctx->poses.n - 1);
    if (P <= 0 || P > P_global) ARGFAIL("bundle_jacobian: bad parameter counts");
    ctx->poses.relative2global();
    CKRC(uploadParams(ctx));
    CKRC(prepareFdBatch(ctx));
    CKRC(transformBase(ctx));
    CKRC(buildSets(ctx, st, sync_build == 0));
    CKRC(runCost(ctx));
    CK(ctx->d_hg.ensure((size_t)P * P + P + 1));
    CK(ctx->d_ls.ensure(16));
    CKRC(jtjInto(ctx, ctx->d_hg.p));
    LAUNCH(k_bundle_scatter, cdiv((size_t)P * P + P + 1, 256), 256, 0, ctx->d_hg.p, P, idx_dev, P_global, ctx->d_linfo.p,
           ctx->levelOn[0] ? ctx->cachedDepth[0] : 1 << 30, ctx->levelOn[1] ? ctx->cachedDepth[1] : 1 << 30, ghg_dev);
    CKRC(ensurePinned(ctx, P));
    CK(cudaMemcpyAsync(pinLinfo(ctx, P), ctx->d_linfo.p, 2 * sizeof(LevelInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}
int dmsa_b200_bundle_line_search(dmsa_b200_ctx* ctx, const double* gstep_dev, const int32_t* idx_dev, double* gls_dev) {
    PdlScope pdl_(ctx);
    if (!gstep_dev || !idx_dev || !gls_dev) ARGFAIL("bundle_line_search: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int P = 6 * (ctx->poses.n - 1);
    LAUNCH(k_bundle_gather_step, cdiv(P, 128), 128, 0, gstep_dev, idx_dev, P, ctx->d_step.p);
    CKRC(lineSearchDev(ctx, ctx->d_ls.p));
    LAUNCH(k_add9, 1, 32, 0, ctx->d_ls.p, gls_dev);
    CK(cudaGetLastError());
    return 0;
}
// after the caller synchronised the stream: the set count of the last bundle_jacobian and whether its deferred build ran on
// a wrong guess (octree depth grew / more sets than the grids were sized for) -> *redo = 1: repeat the iteration with sync_build = 1
int dmsa_b200_bundle_verify(dmsa_b200_ctx* ctx, int32_t* num_gaussians, int32_t* redo) {
    const int P = 6 * (ctx->poses.n - 1);
    if (!ctx->pin) ARGFAIL("bundle_verify: no bundle_jacobian before");
    if (ctx->profiling) profCollect(ctx);
    int G = 0;
    bool miss = false;
    if (ctx->G < 0) {
        memcpy(ctx->h_linfo, pinLinfo(ctx, P), 2 * sizeof(LevelInfo));
        for (int l = 0; l < 2; ++l) {
            if (!ctx->levelOn[l]) continue;
            if (ctx->h_linfo[l].error) ARGFAIL("build_sets: octree deeper than 21 levels (extent / resolution too large)");
            if (ctx->h_linfo[l].depth > ctx->cachedDepth[l]) miss = true;
            ctx->cachedDepth[l] = ctx->h_linfo[l].depth;
            G += ctx->h_linfo[l].G;
        }
        if (G > ctx->cellCap) ARGFAIL("build_sets: Gaussian store capacity exceeded");
        if (G > ctx->Gb) miss = true;
        ctx->Gguess = std::max(G, 1);
        if (!miss) ctx->G = G;
    } else {
        G = ctx->G;
    }
    if (num_gaussians) *num_gaussians = G;
    if (redo) *redo = miss ? 1 : 0;
    return 0;
}
// sum over the ranks of the communicator, in place, on the context's stream (device memory of the caller)
int dmsa_b200_all_reduce(dmsa_b200_ctx* ctx, double* dev, int64_t count) {
    if (!dev || count <= 0) ARGFAIL("all_reduce: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->comm && ctx->world == 1) return 0;
    return allReduceSum(ctx, dev, (size_t)count);
}

// ---- NCCL communicator of a row-sharded context (one process per GPU; NCCL is bound at run time, nccl_dyn.h) ----
int dmsa_b200_comm_unique_id(void* id128) {
    if (!id128) return DMSA_B200_ERR_ARG;
    if (!nccl_api().load()) return DMSA_B200_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) return DMSA_B200_ERR_CUDA;
    memcpy(id128, id.internal, sizeof(id.internal));
    return 0;
}
int dmsa_b200_comm_init(dmsa_b200_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
    if (!id128 || world < 1 || rank < 0 || rank >= world) ARGFAIL("comm_init: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (!nccl_api().load()) {
        ctx->err = "comm_init: " + nccl_api().err;
        return DMSA_B200_ERR_UNSUPPORTED;
    }
    if (ctx->comm) {
        nccl_api().CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    g_commSeen.store(true, std::memory_order_relaxed);
    ncclResult_t r = nccl_api().CommInitRank(&ctx->comm, world, id, rank);
    if (r != ncclSuccess) {
        ctx->comm = nullptr;
        ctx->err = "ncclCommInitRank: " + nccl_api().describe(r);
        return DMSA_B200_ERR_CUDA;
    }
    return dmsa_b200_set_shard(ctx, rank, world);
}
int dmsa_b200_comm_destroy(dmsa_b200_ctx* ctx) {
    if (ctx->comm) {
        CK(cudaStreamSynchronize(ctx->stream));
        nccl_api().CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    return dmsa_b200_set_shard(ctx, 0, 1);
}
int64_t dmsa_b200_collective_count(const dmsa_b200_ctx* ctx) { return ctx ? ctx->collectives : 0; }

int dmsa_b200_cost_jacobian_dev(dmsa_b200_ctx* ctx, double* hg_dev) {
    PdlScope pdl_(ctx);
    CK(cudaSetDevice(ctx->device));
    if (ctx->G <= 0) ARGFAIL("cost_jacobian_dev: build_sets first");
    CKRC(uploadParams(ctx));
    CKRC(prepareFdBatch(ctx));
    CKRC(runCost(ctx));
    return jtjInto(ctx, hg_dev);
}
int dmsa_b200_line_search_costs_dev(dmsa_b200_ctx* ctx, const double* step, double* ls_dev) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->G <= 0) ARGFAIL("line_search_costs_dev: build_sets first");
    CKRC(uploadParams(ctx));
    return lineSearchInto(ctx, step, ls_dev);
}

#include "dmsa_b200_pre.inl"
#include "dmsa_b200_io.inl"

}  // extern "C"
