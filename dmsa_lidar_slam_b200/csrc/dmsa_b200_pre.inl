// dmsa_b200_pre.inl — SURVEY §8(f) rank 3 behind the C-ABI (included by dmsa_b200.cu inside extern "C"):
// randomGridDownsampling (helpers.h:67-182), preProcess (DmsaSlam.h:570-634), updateNormals (DmsaSlam.h:557-568).
// Kernels: kernels_pre.cuh, kernels_knn.cuh; the octree is the set build's (kernels_sets.cuh / kernels_sort.cuh).

}  // extern "C"  (helpers with C++ linkage)
namespace {

// glibc rand() for a given srand() seed (stdlib/random_r.c, TYPE_3: additive feedback generator x^31 + x^3 + 1 seeded by the
// Lehmer generator 16807 mod 2^31 - 1, first 310 outputs discarded, result = top 31 bits) — third-party, restated from the
// published algorithm; tests/test_preprocess_cpu.py checks it against this machine's libc.
void glibcRandSequence(uint32_t seed, size_t n, int32_t* out) {
    std::vector<uint32_t> r(344 + n);
    int32_t word = (int32_t)(seed ? seed : 1u);
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        r[i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (size_t i = 34; i < 344 + n; ++i) r[i] = r[i - 31] + r[i - 3];
    for (size_t k = 0; k < n; ++k) out[k] = (int32_t)(r[k + 344] >> 1);
}

// PCL octree leaves of `n` points at resolution `res` on the device: p_sidx = point indices in leaf (depth-first) order,
// p_raw_start[c] = first slot of leaf c; *R_out = number of leaves.  Two host synchronisations (octree depth, leaf count).
int preLeaves(dmsa_b200_ctx* ctx, const float4* pts, int n, float res, int* R_out) {
    PdlScope pdl_(ctx);
    *R_out = 0;
    if (n <= 0) return 0;
    const int nb = (n + DMSA_KEYS_BLOCK - 1) / DMSA_KEYS_BLOCK;
    const size_t n2 = (size_t)n + 2;
    CK(ctx->p_linfo.ensure(2));
    CK(ctx->p_keys.ensure((size_t)3 * n));
    CK(ctx->p_bb.ensure((size_t)12 * nb));
    CK(ctx->p_code.ensure(n));
    CK(ctx->p_scode.ensure(n));
    CK(ctx->p_idx.ensure(n));
    CK(ctx->p_sidx.ensure(n));
    CK(ctx->p_scan.ensure(n));
    CK(ctx->p_raw_start.ensure(n2));
    CK(ctx->p_raw_diff.ensure(n2));
    cudaStream_t strm = ctx->stream;
    LevelPlan plan;
    plan.n = 1;
    plan.level[0] = 0;
    plan.level[1] = 0;
    plan.res[0] = res;
    plan.res[1] = res;
    LAUNCH(k_anchor, 1, 32, 0, pts, n, plan, ctx->p_linfo.p, n);
    LAUNCH(k_keys, dim3(nb, 1), DMSA_KEYS_BLOCK, 0, pts, n, plan, ctx->p_linfo.p, ctx->p_keys.p, ctx->p_bb.p, nb, ZeroRanges{});
    LAUNCH(k_root, 1, 1024, 0, pts, n, plan, ctx->p_linfo.p, ctx->p_bb.p, nb);
    LevelInfo li;
    CK(cudaMemcpyAsync(&li, ctx->p_linfo.p, sizeof(LevelInfo), cudaMemcpyDeviceToHost, strm));
    CK(cudaStreamSynchronize(strm));
    if (li.error) ARGFAIL("octree deeper than 21 levels (extent / resolution too large)");
    if (li.first >= n) return 0;  // no finite point: no leaf
    const int npass = (std::min(64, 3 * li.depth + 1) + 7) / 8;
    const CtlLayout cl(n, npass, 1);
    CK(ctx->p_ctl.ensure(cl.bytes));
    if (!ctx->sortAttr) {
        CK(cudaFuncSetAttribute(k_sort_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
        ctx->sortAttr = true;
    }
    CK(cudaMemsetAsync(ctx->p_ctl.p, 0, cl.bytes, strm));
    CK(cudaMemsetAsync(ctx->p_raw_diff.p, 0, n2 * sizeof(int), strm));
    SortArgs sa;
    memset(&sa, 0, sizeof(sa));
    const bool odd = (npass & 1) != 0;  // the last pass lands in p_scode / p_sidx
    sa.seg[0].keyA = odd ? ctx->p_code.p : ctx->p_scode.p;
    sa.seg[0].keyB = odd ? ctx->p_scode.p : ctx->p_code.p;
    sa.seg[0].valA = reinterpret_cast<u32_t*>(odd ? ctx->p_idx.p : ctx->p_sidx.p);
    sa.seg[0].valB = reinterpret_cast<u32_t*>(odd ? ctx->p_sidx.p : ctx->p_idx.p);
    sa.nseg = 1;
    sa.n = n;
    sa.tiles = cl.tilesSort;
    sa.npass = npass;
    sa.iota = 1;
    sa.hist = reinterpret_cast<u32_t*>(ctx->p_ctl.p + cl.hist);
    sa.look = reinterpret_cast<u32_t*>(ctx->p_ctl.p + cl.look);
    sa.ticket = reinterpret_cast<int*>(ctx->p_ctl.p + cl.tickets);
    PrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.keys_all = ctx->p_keys.p;
    pa.infos = ctx->p_linfo.p;
    pa.ring = nullptr;  // no ring test here
    pa.flags = ctx->d_flag.p;
    LAUNCH(k_sort_prepare, dim3(cl.tilesSort, 1), RS_T, 0, sa, pa);
    for (int p_ = 0; p_ < npass; ++p_) {
        sa.pass = p_;
        LAUNCH(k_sort_pass, cl.tilesSort, RS_T, RS_SMEM, sa);
    }
    SegmentArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.code[0] = ctx->p_scode.p;
    ga.idx[0] = reinterpret_cast<u32_t*>(ctx->p_sidx.p);
    ga.scan[0] = ctx->p_scan.p;
    ga.raw_start[0] = ctx->p_raw_start.p;
    ga.raw_diff[0] = ctx->p_raw_diff.p;
    ga.infos = ctx->p_linfo.p;
    ga.ring = nullptr;
    ga.flags = ctx->d_flag.p;
    ga.n = n;
    ga.tiles = (n + SG_TILE - 1) / SG_TILE;
    ga.status = reinterpret_cast<u64_t*>(ctx->p_ctl.p + cl.segStatus);
    ga.ticket = reinterpret_cast<int*>(ctx->p_ctl.p + cl.tickets) + 8;
    LAUNCH(k_segment, ga.tiles, SG_T, 0, ga);
    CK(cudaMemcpyAsync(&li, ctx->p_linfo.p, sizeof(LevelInfo), cudaMemcpyDeviceToHost, strm));
    CK(cudaStreamSynchronize(strm));
    CK(cudaGetLastError());
    *R_out = li.R;
    return 0;
}
// randomGridDownsampling: p_pick[c] = index of the member drawn from leaf c (helpers.h:86-106), c = 0 .. R - 1
int preDownsample(dmsa_b200_ctx* ctx, const float4* pts, int n, float grid_size, uint32_t seed, int* R_out) {
    PdlScope pdl_(ctx);
    CKRC(preLeaves(ctx, pts, n, grid_size, R_out));
    const int R = *R_out;
    if (R == 0) return 0;
    std::vector<int32_t> rnd((size_t)R);
    glibcRandSequence(seed, (size_t)R, rnd.data());  // srand(seed); one rand() per leaf, in leaf order
    CK(ctx->p_rand.ensure(R));
    CK(ctx->p_pick.ensure(R));
    CK(cudaMemcpyAsync(ctx->p_rand.p, rnd.data(), (size_t)R * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_pre_pick, cdiv(R, 256), 256, 0, ctx->p_raw_start.p, ctx->p_sidx.p, ctx->p_rand.p, R, ctx->p_pick.p);
    CK(cudaStreamSynchronize(ctx->stream));  // rnd is pageable host memory
    return 0;
}
int preUpload(dmsa_b200_ctx* ctx, const void* points, int64_t n, int stride) {
    CK(ctx->p_raw.ensure((size_t)n * stride));
    CK(ctx->p_pts.ensure((size_t)n));
    CK(cudaMemcpyAsync(ctx->p_raw.p, points, (size_t)n * stride, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_pre_unpack, cdiv(n, 256), 256, 0, ctx->p_raw.p, (int)n, stride, ctx->p_pts.p);
    return 0;
}

}  // namespace
extern "C" {

int dmsa_b200_rand_sequence(uint32_t seed, int64_t n, int32_t* out) {
    if (n < 0 || (n > 0 && !out)) return DMSA_B200_ERR_ARG;
    glibcRandSequence(seed, (size_t)n, out);
    return 0;
}

// randomGridDownsampling(rawPc, filteredPc, gridSize) with srand(seed) (helpers.h:67-182; the reference seeds with time(0)):
// indices_out[c] = index into `points` of the point the reference copies to filteredPc->points[c]; *n_out = leaf count.
int dmsa_b200_grid_downsample(dmsa_b200_ctx* ctx, const void* points, int64_t n, int32_t stride_bytes, float grid_size, uint32_t seed,
                              int32_t* indices_out, int64_t* n_out) {
    if (!n_out || n < 0 || n > 0x3fffffff || (n > 0 && (!points || !indices_out)) || stride_bytes < 16 || stride_bytes % 16 != 0 || !(grid_size > 0.0f))
        ARGFAIL("grid_downsample: bad arguments (stride: a multiple of 16 bytes, xyz floats at offset 0)");
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (n == 0) return 0;
    CKRC(preUpload(ctx, points, n, stride_bytes));
    int R = 0;
    CKRC(preDownsample(ctx, ctx->p_pts.p, (int)n, grid_size, seed, &R));
    if (R > 0) CK(cudaMemcpy(indices_out, ctx->p_pick.p, (size_t)R * sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_out = R;
    return 0;
}
// the same on the staged window's globalPoints (addNewKeyframeToMap, DmsaSlam.h:506)
int dmsa_b200_downsample_global_points(dmsa_b200_ctx* ctx, float grid_size, uint32_t seed, int32_t* indices_out, int64_t* n_out) {
    if (!n_out || !indices_out || !(grid_size > 0.0f)) ARGFAIL("downsample_global_points: bad arguments");
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    const int64_t N = numPoints(ctx);
    if (N == 0) return 0;
    if (!ctx->worldValid) ARGFAIL("downsample_global_points: call update_global_points first");
    int R = 0;
    CKRC(preDownsample(ctx, ctx->d_world.p, (int)N, grid_size, seed, &R));
    if (R > 0) CK(cudaMemcpy(indices_out, ctx->p_pick.p, (size_t)R * sizeof(int32_t), cudaMemcpyDeviceToHost));
    *n_out = R;
    return 0;
}

// preProcess(rawPc, filteredPc) (DmsaSlam.h:570-634) with srand(seed) in every randomGridDownsampling call.
// out: room for n records; *n_out records are written; *grid_size_out = filteredPc->gridSize.
int dmsa_b200_preprocess_scan(dmsa_b200_ctx* ctx, const dmsa_b200_point_stamp_id* raw, int64_t n, const dmsa_b200_preprocess_config* cfg, uint32_t seed,
                              dmsa_b200_point_stamp_id* out, int64_t* n_out, float* grid_size_out) {
    PdlScope pdl_(ctx);
    if (!cfg || !n_out || n < 0 || n > 0x3fffffff || (n > 0 && (!raw || !out))) ARGFAIL("preprocess_scan: bad arguments");
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    if (grid_size_out) *grid_size_out = 0.4f;
    if (n == 0) return 0;
    const int stride = (int)sizeof(dmsa_b200_point_stamp_id);
    CKRC(preUpload(ctx, raw, n, stride));
    // adaptive random grid filter (:573-593): the next finer grid while the cloud has fewer than max_num points
    const float grids[4] = {0.4f, 0.3f, 0.2f, 0.15f};
    int R = 0;
    float used = grids[0];
    for (int k = 0; k < 4; ++k) {
        if (k > 0 && !((size_t)R < (size_t)cfg->max_num_points_per_scan)) break;  // size_t comparison like the reference's
        used = grids[k];
        CKRC(preLeaves(ctx, ctx->p_pts.p, (int)n, used, &R));
    }
    if (grid_size_out) *grid_size_out = used;
    if (R == 0) return 0;
    {
        std::vector<int32_t> rnd((size_t)R);
        glibcRandSequence(seed, (size_t)R, rnd.data());
        CK(ctx->p_rand.ensure(R));
        CK(ctx->p_pick.ensure(R));
        CK(cudaMemcpyAsync(ctx->p_rand.p, rnd.data(), (size_t)R * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_pre_pick, cdiv(R, 256), 256, 0, ctx->p_raw_start.p, ctx->p_sidx.p, ctx->p_rand.p, R, ctx->p_pick.p);
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // ranges, their sorted copy (:596-606), threshold and the kept points (:609-623)
    CK(ctx->p_range.ensure(R));
    CK(ctx->p_flag.ensure((size_t)R + 1));
    CK(ctx->p_pos.ensure((size_t)R + 1));
    CK(ctx->p_out.ensure((size_t)R * stride));
    // (the octree buffers are free again: p_code / p_scode hold the range keys, p_idx / p_sidx the sort's values)
    LAUNCH(k_pre_ranges, cdiv(R, 256), 256, 0, ctx->p_pts.p, ctx->p_pick.p, R, ctx->p_range.p, ctx->p_code.p);
    const int npass = 4;  // 32-bit keys
    const CtlLayout cl(R, npass, 1);
    CK(ctx->p_ctl.ensure(cl.bytes));
    CK(cudaMemsetAsync(ctx->p_ctl.p, 0, cl.bytes, ctx->stream));
    SortArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.seg[0].keyA = ctx->p_code.p;  // even number of passes: the result is back in keyA
    sa.seg[0].keyB = ctx->p_scode.p;
    sa.seg[0].valA = reinterpret_cast<u32_t*>(ctx->p_idx.p);
    sa.seg[0].valB = reinterpret_cast<u32_t*>(ctx->p_sidx.p);
    sa.nseg = 1;
    sa.n = R;
    sa.tiles = cl.tilesSort;
    sa.npass = npass;
    sa.iota = 1;
    sa.hist = reinterpret_cast<u32_t*>(ctx->p_ctl.p + cl.hist);
    sa.look = reinterpret_cast<u32_t*>(ctx->p_ctl.p + cl.look);
    sa.ticket = reinterpret_cast<int*>(ctx->p_ctl.p + cl.tickets);
    LAUNCH(k_sort_hist, dim3(cl.tilesSort, 1), RS_T, 0, sa);
    for (int p_ = 0; p_ < npass; ++p_) {
        sa.pass = p_;
        LAUNCH(k_sort_pass, cl.tilesSort, RS_T, RS_SMEM, sa);
    }
    LAUNCH(k_pre_keep, cdiv(R, 256), 256, 0, ctx->p_range.p, ctx->p_code.p, R, cfg->max_num_points_per_scan, cfg->min_dist_ds, cfg->min_dist, ctx->p_flag.p);
    {
        ScanArgs sc;
        sc.in = ctx->p_flag.p;
        sc.out = ctx->p_pos.p;
        sc.n = R;
        sc.tiles = (R + CS_TILE - 1) / CS_TILE;
        sc.status = reinterpret_cast<u64_t*>(ctx->p_ctl.p + cl.scanStatus[0]);
        sc.ticket = reinterpret_cast<int*>(ctx->p_ctl.p + cl.tickets) + 10;
        LAUNCH(k_scan_excl, sc.tiles, CS_T, 0, sc);
    }
    PreTform T;
    for (int q = 0; q < 16; ++q) T.m[q] = cfg->lidar_to_imu[q];
    CK(ctx->d_gcount.ensure(1));
    LAUNCH(k_pre_emit, cdiv(R, 256), 256, 0, ctx->p_raw.p, stride, ctx->p_pick.p, ctx->p_flag.p, ctx->p_pos.p, R, T, ctx->p_out.p, ctx->d_gcount.p);
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->d_gcount.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (cnt > 0) CK(cudaMemcpy(out, ctx->p_out.p, (size_t)cnt * stride, cudaMemcpyDeviceToHost));
    *n_out = cnt;
    return 0;
}

// updateNormals(cloud, origin) (DmsaSlam.h:557-568): pcl::NormalEstimationOMP, setKSearch(6), search surface = the cloud,
// normals flipped towards `viewpoint`; normal_x/y/z and curvature of every point are overwritten in place.
// cell_size: edge of the search grid (> 0; the cloud's grid size is a good value: about one point per cell).
// nn_indices (optional): the 6 neighbour indices of every point in search-result order, -1 where fewer exist.
int dmsa_b200_estimate_normals(dmsa_b200_ctx* ctx, dmsa_b200_point_normal* cloud, int64_t n, const float* viewpoint, float cell_size, int32_t* nn_indices) {
    PdlScope pdl_(ctx);
    if (n < 0 || n > 0x3fffffff || (n > 0 && !cloud) || !viewpoint || !(cell_size > 0.0f)) ARGFAIL("estimate_normals: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    CK(ctx->p_cloud.ensure((size_t)3 * n));
    CK(cudaMemcpyAsync(ctx->p_cloud.p, cloud, (size_t)n * 48, cudaMemcpyHostToDevice, ctx->stream));
    HashGrid g;
    ctx->gridEpoch = ~0ull;  // the grid buffers now hold this cloud, not the window
    CKRC(buildHashGrid(ctx, ctx->p_cloud.p, (int)n, 3, (double)cell_size, &g));
    if (nn_indices) CK(ctx->p_nn.ensure((size_t)KNN_K * n));
    LAUNCH(k_normals_knn6, cdiv(n, 128), 128, 0, g, ctx->p_cloud.p, (int)n, viewpoint[0], viewpoint[1], viewpoint[2], nn_indices ? ctx->p_nn.p : nullptr);
    CK(cudaMemcpyAsync(cloud, ctx->p_cloud.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
    if (nn_indices) CK(cudaMemcpyAsync(nn_indices, ctx->p_nn.p, (size_t)KNN_K * n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}
