// internal.h -- library-internal seams between the translation units of libmambo_b200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include "../../include/mambo_csa.h"

// read-only view of the resident residual of an unsharded handle (df.cu)
struct mambo_resident_view {
    int N, mode, packed, rows, device;
    double asym_pair;               // measured relative asymmetry under (pq) <-> (rs) of the tensor given to create
    long long ldT;
    const double *T1, *T2;          // T~ (and its transpose in GENERAL mode), row-major, pitch ldT
    const int *pa, *pq;             // row -> (p, q)
    const double* sw;               // sqrt(weight) of the packed rows
    cudaStream_t stream;
};
extern "C" int mambo_csa_resident_view(mambo_csa* h, mambo_resident_view* v);
// one opaque, handle-owned state object per handle (the double factorisation); freed by mambo_csa_destroy through `destroy`
extern "C" void** mambo_csa_df_slot(mambo_csa* h, void (*destroy)(void*));
// the handle's own device-resident x (num_params) and [cost, grad] (1 + num_params) buffers: an evaluation called with exactly these
// pointers skips its two device-to-device copies (optimizer.cu works in them)
extern "C" int mambo_csa_io_buffers(mambo_csa* h, double** d_x, double** d_out);
extern "C" void mambo_set_error(const char* msg);
extern "C" void mambo_csa_count_launches(mambo_csa* h, long long n);
// 1 when the handle evaluates through the single-kernel path (tiny::k_eval): one evaluation is one plain kernel launch, so the
// optimiser may capture a whole device-driven iteration into a CUDA graph
extern "C" int mambo_csa_is_tiny(mambo_csa* h);
// mambo_csa_eval_device without the argument checks and cudaSetDevice (safe inside a stream capture on the handle's stream)
extern "C" int mambo_csa_enqueue_eval(mambo_csa* h, const double* d_x, double* d_out);
