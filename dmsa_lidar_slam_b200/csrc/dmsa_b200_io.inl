// dmsa_b200_io.inl — SURVEY §8(f) rank 4: the data formats either side of the path (included by dmsa_b200.cu inside extern "C").
//   src/dmsa_slam_ros.cpp:372-486  sensor_msgs/PointCloud2 byte buffer -> PointStampId records (one branch per sensor type)
//   OutputManagement.h:80-96       addPoseToFile: one TUM trajectory line (stamp, translation, quaternion of axang2rotm)
//   src/dmsa_slam_ros.cpp:286-291  pcl::io::savePCDFileASCII of the keyframe map (pcl::PointNormal)

}  // extern "C"
namespace {

// One thread per point: byte-wise gathers out of the message buffer (fields may be unaligned), one 32-byte record out.
__global__ void k_decode_pc2(const unsigned char* __restrict__ data, int n, dmsa_b200_pc2_layout L, double stamp_msg, double delta_t,
                             dmsa_b200_point_stamp_id* __restrict__ out) {
    DMSA_PDL_ENTER();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned char* p = data + (size_t)k * L.point_step;
    auto rd = [&](int off, void* dst, int bytes) {
        unsigned char* d = reinterpret_cast<unsigned char*>(dst);
        for (int b = 0; b < bytes; ++b) d[b] = p[off + b];
    };
    dmsa_b200_point_stamp_id o;
    rd(L.x_offset, &o.x, 4);
    rd(L.y_offset, &o.y, 4);
    rd(L.z_offset, &o.z, 4);
    o.w = 1.0f;  // PointStampId's PCL_ADD_POINT4D constructor sets data[3] = 1
    double stamp = 0.0;
    switch (L.stamp_type) {
        case DMSA_B200_STAMP_F64_ABS: {  // hesai / robosense / livoxXYZRTLT_s
            double t;
            rd(L.stamp_offset, &t, 8);
            stamp = t;
        } break;
        case DMSA_B200_STAMP_U32_NS_REL: {  // ouster: stampMsg + 1e-9 * (double) relStampNano
            unsigned t;
            rd(L.stamp_offset, &t, 4);
            stamp = stamp_msg + 1e-9 * (double)t;
        } break;
        case DMSA_B200_STAMP_F32_REL: {  // velodyne / sick: stampMsg + (double) float
            float t;
            rd(L.stamp_offset, &t, 4);
            stamp = stamp_msg + (double)t;
        } break;
        case DMSA_B200_STAMP_F64_NS_ABS: {  // livoxXYZRTLT_ns: 1e-9 * double
            double t;
            rd(L.stamp_offset, &t, 8);
            stamp = 1e-9 * t;
        } break;
        default:  // unknown sensor: stampMsg + deltaTPcs * (double) k / (double) (height * width)
            stamp = stamp_msg + delta_t * (double)k / (double)(unsigned)n;
    }
    int id;
    switch (L.ring_type) {
        case DMSA_B200_RING_U16: {
            unsigned short r;
            rd(L.ring_offset, &r, 2);
            id = (int)r;
        } break;
        case DMSA_B200_RING_U8: id = (int)p[L.ring_offset]; break;
        case DMSA_B200_RING_I8: id = (int)(signed char)p[L.ring_offset]; break;
        default: id = k % 1000;  // artificial ring index
    }
    o.stamp = stamp;
    o.id = id;
    o.isStatic = 0;
    out[k] = o;
}

// Eigen::Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl<.., 3, 3>): x, y, z, w
void quatFromMat(const Mat3& R, double q[4]) {
    const double* m = R.m;  // row-major
    double t = m[0] + m[4] + m[8];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

}  // namespace
extern "C" {

// callbackPointCloud's per-point loop (src/dmsa_slam_ros.cpp:399-483) on the device: data = msg->data (n_points * point_step
// bytes on the host), out = n_points PointStampId records on the host.
int dmsa_b200_decode_pointcloud2(dmsa_b200_ctx* ctx, const uint8_t* data, int64_t n_points, const dmsa_b200_pc2_layout* layout, double stamp_msg,
                                 double delta_t, dmsa_b200_point_stamp_id* out) {
    if (!layout || n_points < 0 || n_points > 0x3fffffff || (n_points > 0 && (!data || !out)) || layout->point_step <= 0)
        ARGFAIL("decode_pointcloud2: bad arguments");
    const int ps = layout->point_step;
    auto inside = [&](int off, int bytes) { return off >= 0 && off + bytes <= ps; };
    const int sb = layout->stamp_type == DMSA_B200_STAMP_NONE ? 0 : ((layout->stamp_type == DMSA_B200_STAMP_F64_ABS || layout->stamp_type == DMSA_B200_STAMP_F64_NS_ABS) ? 8 : 4);
    const int rb = layout->ring_type == DMSA_B200_RING_NONE ? 0 : (layout->ring_type == DMSA_B200_RING_U16 ? 2 : 1);
    if (!inside(layout->x_offset, 4) || !inside(layout->y_offset, 4) || !inside(layout->z_offset, 4) || (sb && !inside(layout->stamp_offset, sb)) ||
        (rb && !inside(layout->ring_offset, rb)) || layout->stamp_type < 0 || layout->stamp_type > 4 || layout->ring_type < 0 || layout->ring_type > 3)
        ARGFAIL("decode_pointcloud2: a field lies outside the point record or has an unknown type");
    CK(cudaSetDevice(ctx->device));
    if (n_points == 0) return 0;
    CK(ctx->p_raw.ensure((size_t)n_points * ps));
    CK(ctx->p_out.ensure((size_t)n_points * sizeof(dmsa_b200_point_stamp_id)));
    CK(cudaMemcpyAsync(ctx->p_raw.p, data, (size_t)n_points * ps, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_decode_pc2, cdiv(n_points, 256), 256, 0, ctx->p_raw.p, (int)n_points, *layout, stamp_msg, delta_t, reinterpret_cast<dmsa_b200_point_stamp_id*>(ctx->p_out.p));
    CK(cudaMemcpyAsync(out, ctx->p_out.p, (size_t)n_points * sizeof(dmsa_b200_point_stamp_id), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}
// the field layouts the reference hard-codes per sensor type (offsets of msg->fields[i] are the caller's: fields[] = offsets by index)
int dmsa_b200_pc2_layout_for_sensor(const char* sensor, const int32_t* field_offsets, int32_t n_fields, int32_t point_step, dmsa_b200_pc2_layout* out) {
    if (!sensor || !field_offsets || !out || n_fields < 3) return DMSA_B200_ERR_ARG;
    auto f = [&](int i) { return i < n_fields ? field_offsets[i] : -1; };
    dmsa_b200_pc2_layout L;
    L.point_step = point_step;
    L.x_offset = f(0);
    L.y_offset = f(1);
    L.z_offset = f(2);
    L.stamp_offset = L.ring_offset = 0;
    L.stamp_type = DMSA_B200_STAMP_NONE;
    L.ring_type = DMSA_B200_RING_NONE;
    const std::string s(sensor);
    if (s == "hesai") { L.stamp_offset = f(4); L.stamp_type = DMSA_B200_STAMP_F64_ABS; L.ring_offset = f(5); L.ring_type = DMSA_B200_RING_U16; }
    else if (s == "ouster") { L.stamp_offset = f(4); L.stamp_type = DMSA_B200_STAMP_U32_NS_REL; L.ring_offset = f(6); L.ring_type = DMSA_B200_RING_U8; }
    else if (s == "robosense") { L.stamp_offset = f(5); L.stamp_type = DMSA_B200_STAMP_F64_ABS; L.ring_offset = f(4); L.ring_type = DMSA_B200_RING_U16; }
    else if (s == "velodyne") { L.stamp_offset = f(5); L.stamp_type = DMSA_B200_STAMP_F32_REL; L.ring_offset = f(4); L.ring_type = DMSA_B200_RING_U16; }
    else if (s == "livoxXYZRTLT_s") { L.stamp_offset = f(6); L.stamp_type = DMSA_B200_STAMP_F64_ABS; }
    else if (s == "livoxXYZRTLT_ns") { L.stamp_offset = f(6); L.stamp_type = DMSA_B200_STAMP_F64_NS_ABS; }
    else if (s == "sick") { L.stamp_offset = f(8); L.stamp_type = DMSA_B200_STAMP_F32_REL; L.ring_offset = f(11); L.ring_type = DMSA_B200_RING_I8; }
    else if (s != "unknown") return DMSA_B200_ERR_ARG;
    if ((L.stamp_type != DMSA_B200_STAMP_NONE && L.stamp_offset < 0) || (L.ring_type != DMSA_B200_RING_NONE && L.ring_offset < 0)) return DMSA_B200_ERR_ARG;
    *out = L;
    return 0;
}

// addPoseToFile (OutputManagement.h:80-96): "stamp x y z qx qy qz qw\n" with the reference's precisions; returns the
// number of characters written (excluding the terminating NUL), or -1 if buf is too small.
int dmsa_b200_format_tum_pose(double stamp, const double* pos, const double* orient, char* buf, int32_t buf_size) {
    if (!pos || !orient || !buf || buf_size <= 0) return -1;
    const Vec3 aa = mk3(orient[0], orient[1], orient[2]);
    const Mat3 R = so3_exp(aa);  // axang2rotm (helpers.h:51-57, identity below 1e-5 rad)
    double q[4];
    quatFromMat(R, q);
    const int n = snprintf(buf, (size_t)buf_size, "%.6f %.5f %.5f %.5f %.6f %.6f %.6f %.6f\n", stamp, pos[0], pos[1], pos[2], q[0], q[1], q[2], q[3]);
    return (n < 0 || n >= buf_size) ? -1 : n;
}

// pcl::io::savePCDFileASCII(filename, PointCloud<PointNormal>) (PCL 1.10 io/pcd_io.h writeASCII, precision 8): header +
// one line per point "x y z normal_x normal_y normal_z curvature"; returns 0, or -1 if the file cannot be written.
int dmsa_b200_save_pcd_ascii(const char* filename, const dmsa_b200_point_normal* cloud, int64_t n) {
    if (!filename || n < 0 || (n > 0 && !cloud)) return -1;
    FILE* f = fopen(filename, "w");
    if (!f) return -1;
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z normal_x normal_y normal_z curvature\nSIZE 4 4 4 4 4 4 4\n"
               "TYPE F F F F F F F\nCOUNT 1 1 1 1 1 1 1\nWIDTH %lld\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %lld\nDATA ascii\n",
            (long long)n, (long long)n);
    for (int64_t i = 0; i < n; ++i) {
        const float v[7] = {cloud[i].x, cloud[i].y, cloud[i].z, cloud[i].nx, cloud[i].ny, cloud[i].nz, cloud[i].curvature};
        for (int k = 0; k < 7; ++k) {
            if (std::isnan(v[k]))
                fputs("nan", f);  // PCL writes NaN as "nan"
            else
                fprintf(f, "%.8g", (double)v[k]);  // std::ostream << float with precision(8), classic locale
            fputc(k == 6 ? '\n' : ' ', f);
        }
    }
    return fclose(f) == 0 ? 0 : -1;
}

// ConsecutivePoses.h:26-43 relative2global on plain arrays (3 x n column-major doubles; host only, no context): the pose chain
// the bundle driver needs once per iteration to place every bundle's first keyframe.
int dmsa_b200_relative2global(int32_t n, const double* rel_orient, const double* rel_transl, double* glob_orient, double* glob_transl) {
    if (n <= 0 || !rel_orient || !rel_transl || !glob_orient || !glob_transl) return DMSA_B200_ERR_ARG;
    HostPoses hp;
    hp.resize(n);
    std::copy(rel_orient, rel_orient + 3 * (size_t)n, hp.relO.begin());
    std::copy(rel_transl, rel_transl + 3 * (size_t)n, hp.relT.begin());
    hp.relative2global();
    std::copy(hp.globO.begin(), hp.globO.end(), glob_orient);
    std::copy(hp.globT.begin(), hp.globT.end(), glob_transl);
    return 0;
}
