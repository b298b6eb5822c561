// tcgen05 3x3 convolutions (halo-reuse + TMA implicit GEMM), bf16 operands
#define EDV_GEMM_TU_CONV
#include "gemm_tc_inst.cuh"
namespace edv { template void launch_gemm_tc_conv<bf16>(Launch&, int, const GemmArgs&); }
