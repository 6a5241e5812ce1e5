// Oracle shim: Sophus is absent; common_include.h only typedefs Sophus::SE3d / SO3d (never used by the extractor).
#ifndef SLAMB200_ORACLE_SHIM_SOPHUS_SE3
#define SLAMB200_ORACLE_SHIM_SOPHUS_SE3
namespace Sophus {
template <typename Scalar, int Options = 0> class SE3;
template <typename Scalar, int Options = 0> class SO3;
typedef SE3<double> SE3d;
typedef SO3<double> SO3d;
}
#endif
