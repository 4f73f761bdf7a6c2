// TEST INFRASTRUCTURE: stands in for upstream Kaldi cudamatrix/cu-matrixdim.h so that the reference's bd-cu-kernels.cu compiles under nvcc
#ifndef ORACLE_CU_MATRIXDIM_H_
#define ORACLE_CU_MATRIXDIM_H_
#include <stdint.h>
#ifndef HAVE_CUDA
#define HAVE_CUDA 1
#endif
extern "C" {
typedef struct MatrixDim_ {
  int32_t rows, cols, stride;
} MatrixDim;
typedef int32_t int32_cuda;
}
#endif
