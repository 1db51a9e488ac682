// reference_mock.h — mock classes carrying the PUBLIC MEMBER NAMES of the reference's types that the reference-types
// binding (DmsaOptimizerB200T, dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h) touches:
//   OptimizablePointSet<PointT>   OptimizablePointSet.h:18-56     (globalPoints, minGridSize)
//   ContinuousTrajectory          ContinuousTrajectory.h:24-72    (controlPoses, denseGlobalPoses, denseTformsLocal2Global,
//                                                                  gravity, preint*, CovPVRot_inv, t0, horizon, dt_res, ...)
//   MapManagement / KeyframeData  MapManagement.h:20-70, KeyframeData.h:17-33
//   Poses / StampedConsecutivePoses, PointCloudBuffer / PointCloudPlus / RingBuffer, PointStampId, pcl::PointNormal,
//   DmsaOptimSettings (DmsaOptimizer.h:25-39)
// Eigen and PCL are absent from this image, so the Eigen objects are minimal column-major stand-ins offering data() and
// operator() only — which is all the binding uses.  TEST INFRASTRUCTURE: lets the guarded branch compile and run here.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace mock_eigen {
template <class T, int R, int C>
struct Fixed {  // column-major like Eigen's default
    T v[R * C] = {};
    T* data() { return v; }
    const T* data() const { return v; }
    T& operator()(int r, int c) { return v[c * R + r]; }
    const T& operator()(int r, int c) const { return v[c * R + r]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
};
template <class T, int R>
struct DynCols {  // R x n, column-major
    std::vector<T> v;
    int n = 0;
    void resize(int n_) {
        n = n_;
        v.assign((size_t)R * n, T(0));
    }
    T* data() { return v.data(); }
    int cols() const { return n; }
    T& operator()(int r, int c) { return v[(size_t)c * R + r]; }
};
template <class T>
struct DynVec {
    std::vector<T> v;
    void resize(int n) { v.assign(n, T(0)); }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    int size() const { return (int)v.size(); }
    T& operator()(int i) { return v[i]; }
};
}  // namespace mock_eigen
using Matrix3d = mock_eigen::Fixed<double, 3, 3>;
using Vector3d = mock_eigen::Fixed<double, 3, 1>;
using Matrix4f = mock_eigen::Fixed<float, 4, 4>;
using Matrix3Xd = mock_eigen::DynCols<double, 3>;
using VectorXd = mock_eigen::DynVec<double>;
using VectorXi = mock_eigen::DynVec<int>;
template <class T, int R, int C>
using Matrix = mock_eigen::Fixed<T, R, C>;

// PointStampId.h:33-45
struct alignas(16) PointStampId {
    union {
        float data[4];
        struct {
            float x, y, z;
        };
    };
    double stamp;
    int id;
    int isStatic;
};
namespace pcl {
struct alignas(16) PointNormal {
    union {
        float data[4];
        struct {
            float x, y, z;
        };
    };
    union {
        float data_n[4];
        struct {
            float normal_x, normal_y, normal_z;
        };
    };
    float curvature;
    float pad_[3];
};
template <class PointT>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    std::vector<PointT> points;
    size_t size() const { return points.size(); }
    void resize(size_t n) { points.resize(n); }
};
}  // namespace pcl

// DmsaOptimizer.h:25-39
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

// OptimizablePointSet.h:18-56 (the members the binding reads / writes)
template <typename PointT>
class OptimizablePointSet {
public:
    pcl::PointCloud<PointT> globalPoints;
    float minGridSize = 0.3;
    virtual ~OptimizablePointSet() {}
};

// Poses.h:16-20, ConsecutivePoses.h
struct Poses {
    Matrix3Xd Orientations, Translations;
    int numPoses = 0;
    void resize(int n) {
        numPoses = n;
        Orientations.resize(n);
        Translations.resize(n);
    }
};
struct StampedConsecutivePoses {
    Poses relativePoses, globalPoses;
    VectorXd stamps;
    int numPoses = 0;
    StampedConsecutivePoses() {}
    explicit StampedConsecutivePoses(int n) : numPoses(n) {
        relativePoses.resize(n);
        globalPoses.resize(n);
        stamps.resize(n);
    }
};

// PointCloudPlus.h:15-18, PointCloudBuffer.h, RingBuffer.h:31-52
struct PointCloudPlus : public pcl::PointCloud<PointStampId> {
    float gridSize = 0.3f;
};
template <class T>
struct RingBuffer {
    std::vector<T> elems;
    void init(int n) { elems.reserve(n); }
    void addElem(const T& e) { elems.push_back(e); }
    T& at(int chronologicalIndex) { return elems[chronologicalIndex]; }
    int getNumElements() { return (int)elems.size(); }
};
struct PointCloudBuffer : public RingBuffer<PointCloudPlus> {
    int getNumPoints() {
        int n = 0;
        for (auto& e : elems) n += (int)e.size();
        return n;
    }
};

// ContinuousTrajectory.h:24-72
class ContinuousTrajectory : public OptimizablePointSet<PointStampId> {
public:
    StampedConsecutivePoses controlPoses;
    Poses denseGlobalPoses;
    std::vector<Matrix4f> denseTformsLocal2Global;
    VectorXd trajTime;
    Vector3d gravity;
    std::vector<Matrix3d> preintImuRots;
    std::vector<Vector3d> preintRelPositions;
    std::vector<Vector3d> preintRelVelocity;
    std::vector<Matrix<double, 9, 9>> CovPVRot_inv;
    double t0 = 0;
    double horizon = 0;
    double dt_res = 0.0001;
    double balancingImu = 0.001f;
    bool useImuErrorTerms = false;
    int numParams = 0;
    int n_total = 0;
    std::shared_ptr<PointCloudBuffer> regPcBuffer;
};

// KeyframeData.h:17-33
class KeyframeData {
public:
    pcl::PointCloud<pcl::PointNormal>::Ptr pointCloudLocal;
    VectorXi ringIds;
    float gridSize = 0.3f;
    Vector3d measuredGravity;
    bool gravityPlausible = false;
    Vector3d relativeTransl;
    Vector3d relativeOrient;
    Matrix3d relativeOrientMat;
};
// MapManagement.h:20-70
class MapManagement : public OptimizablePointSet<pcl::PointNormal> {
public:
    StampedConsecutivePoses keyframePoses;
    RingBuffer<KeyframeData> keyframeDataBuffer;
    bool useGravityErrorTerms = false;
    bool useOdometryErrorTerms = false;
    Vector3d gravity;
    double std_dev_acc = 0.3;
    Matrix3d odometryTranslCovInv;
    Matrix3d odometryOrientCovInv;
    Matrix3d Cov_grav_inv;
    double balancingFactorGrav = 1.0;
    double balancingFactorOdom = 1000.0;
    explicit MapManagement(int n_max = 30) : keyframePoses(n_max) {
        keyframeDataBuffer.init(n_max);
        gravity(0) = 0.0;
        gravity(1) = 0.0;
        gravity(2) = -9.805;
        for (int i = 0; i < 3; ++i) {  // MapManagement.h:66-70 (inverse of a scaled identity)
            Cov_grav_inv(i, i) = 1.0 / (std_dev_acc * std_dev_acc);
            odometryTranslCovInv(i, i) = 1.0 / (0.01 * 0.01);
            odometryOrientCovInv(i, i) = 1.0 / (0.01 * 0.01);
        }
    }
};
