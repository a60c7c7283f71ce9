// l <= 2 tensor product with a scalar second input as plain per-irrep linear maps (o3tp_lin.cu): host interface of o3tp.cu.
#pragma once
#include <cuda_runtime.h>

#include "o3tp_tables.h"

struct O3Lin;
// nullptr unless in2 is one l = 0 irrep and every in1 / out irrep has a path (the node tables of se3gnn_b200/o3msg.py)
O3Lin* o3lin_create(const o3::Plan& P);
void o3lin_destroy(O3Lin* s);
int o3lin_forward(O3Lin* s, long long rows, const float* x, const float* y, const float* w, float* out, cudaStream_t st);
// gx [rows, d_in1] (overwritten) from the cotangent g [rows, d_out]; dense rows
int o3lin_gin(O3Lin* s, long long rows, const float* g, const float* y, const float* w, float* gx, cudaStream_t st);
