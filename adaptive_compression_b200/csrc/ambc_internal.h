// ambc_internal.h -- host-side plumbing shared by the .cu files of libambc.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../include/ambc.h"

using std::min;
using std::max;

int ambc_fail(int code, const char *fmt, ...);
void ambc_count_launch();

#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return ambc_fail(AMBC_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

// optional per-kernel timing (ambc_enable_timing)
struct AmbcTiming {
    bool on = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float ms[4] = {0, 0, 0, 0};
    bool pending_c = false, pending_d = false;
};
AmbcTiming &ambc_timing();
void ambc_timing_mark(int idx, cudaStream_t s);

// host-buffer compress: where the body goes and what the piece-wise download needs
struct AmbcPieceOut {
    void *out_host;        // destination of the body (pinned memory recommended)
    uint64_t out_cap;
    void *states;          // pinned, n_pieces * ambc_scan_state_bytes()
    cudaEvent_t *done;     // one event per piece
    cudaStream_t d2h;      // download stream
    cudaStream_t aux;      // scan / pack stream (nullptr: same stream as k_select)
    cudaEvent_t *sel;      // one event per piece: its k_select finished
};
extern "C" uint64_t ambc_scan_state_bytes(void);
