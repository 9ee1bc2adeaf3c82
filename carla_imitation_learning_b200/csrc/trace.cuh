// CTA timeline for -DBC_TRACE builds (tools/*_trace.py): every translation unit that includes this header gets its OWN buffer and
// exports its own getter through BC_TRACE_EXPORT(name); the shipped library compiles all of it away.
//   TRACE_DECL            once per kernel, after `warp` / `lane` exist and tr_t0 has been taken (TRACE_T0)
//   TRACE(event, id)      lane 0 of every warp of the traced CTA appends (event << 56 | id << 40 | cycles since TRACE_T0)
//   TRACE_END             terminates the warp's list
#pragma once
#include <cuda_runtime.h>
#ifdef BC_TRACE
static __device__ unsigned long long g_trace[20][1024];
static __device__ int g_trace_cta = 0;
#define TRACE_T0 const long long tr_t0 = clock64();
#define TRACE_DECL uint32_t tr_n = 0; const bool tr_on = (int)blockIdx.x == g_trace_cta && (int)blockIdx.y == 0 && lane == 0;
#define TRACE(ev, id) do { if (tr_on && tr_n < 1023) g_trace[warp][tr_n++] = ((unsigned long long)(ev) << 56) | ((unsigned long long)((id) & 0xffff) << 40) | (unsigned long long)((clock64() - tr_t0) & 0xffffffffffull); } while (0)
#define TRACE_END do { if (tr_on) g_trace[warp][tr_n] = ~0ull; } while (0)
#define BC_TRACE_EXPORT(name) extern "C" int name(unsigned long long* host_out, int cta) { \
    if (host_out == nullptr) return cudaMemcpyToSymbol(g_trace_cta, &cta, sizeof(int)) == cudaSuccess ? 0 : -1; \
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(g_trace)) == cudaSuccess ? 0 : -1; }
#else
#define TRACE_T0
#define TRACE_DECL
#define TRACE(ev, id) do {} while (0)
#define TRACE_END do {} while (0)
#define BC_TRACE_EXPORT(name)
#endif
