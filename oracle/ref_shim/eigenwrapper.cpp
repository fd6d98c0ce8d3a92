#include <vector>
#include "eigenwrapper.h"  // the shim header next to this file

namespace auxil {
template <class T> T* MakeIMatrix(int height, int width) {
    static std::vector<T> m;
    m.assign((size_t)height * width, T(0));
    for (int i = 0; i < height && i < width; ++i) m[(size_t)i * height + i] = T(1);
    return m.data();
}
template <class T> T* MakeZMatrix(int height, int width) {
    static std::vector<T> m;
    m.assign((size_t)height * width, T(0));
    return m.data();
}
double* MakeFMatrix(int height, int width, double SampleLength) {
    if (height < 8 || width < 8) return nullptr;
    static std::vector<double> m;
    m.assign((size_t)height * width, 0.0);
    for (int i = 0; i < height && i < width; ++i) m[(size_t)i * height + i] = 1.0;
    for (int i = 0; i < 4; ++i) m[(size_t)(i + 4) * height + i] = SampleLength;   // block(0,4), column-major
    return m.data();
}
template double* MakeIMatrix<double>(int, int);
template float* MakeIMatrix<float>(int, int);
template double* MakeZMatrix<double>(int, int);
template float* MakeZMatrix<float>(int, int);
}
