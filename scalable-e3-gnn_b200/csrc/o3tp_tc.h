// Tensor-core (tcgen05) kernels of the l <= 2 tensor product: host interface used by o3tp.cu.
#pragma once
#include <cuda_runtime.h>

#include "o3tp_tables.h"

struct O3TcGw;   // per-plan state of the weight-gradient kernel (o3tp_tc_gw.cu)

// nullptr if the plan is not covered (d_in1 > 64, more than 512 GT columns, ...): the caller keeps the SIMT kernel
O3TcGw* o3tp_tc_gw_create(const o3::Plan& P);
void o3tp_tc_gw_destroy(O3TcGw* s);
bool o3tp_tc_gw_aligned(const O3TcGw* s, const float* x, const float* y, const float* g);
// gw (overwritten) = weight gradient over the first `rows` rows, rows a positive multiple of 32; in1 dense [rows, d_in1]
int o3tp_tc_gw_run(O3TcGw* s, long long rows, const float* x, const float* y, const float* g, float* gw, cudaStream_t st);
