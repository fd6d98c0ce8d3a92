/* Shim placed first on the include path when building the reference for the
 * oracle (oracle/Makefile): the reference's auxil/inc/eigenwrapper.h pulls in
 * Eigen, which this image does not have, and uses it only to produce identity /
 * zero / F matrices (auxil/src/eigenwrapper.cpp:11-46).  Same three entry
 * points, std::vector storage, column-major like Eigen's default. */
#ifndef ORACLE_REF_SHIM_EIGENWRAPPER_H_
#define ORACLE_REF_SHIM_EIGENWRAPPER_H_
namespace auxil {
template <class T> T* MakeIMatrix(int height, int width);
template <class T> T* MakeZMatrix(int height, int width);
double* MakeFMatrix(int height, int width, double SampleLength);
}
#endif
