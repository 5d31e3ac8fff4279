// TEST INFRASTRUCTURE — a host-side stand-in for <cuda_runtime.h> that lets the product's kernel headers
// (word2bits_b200/csrc/*.cuh) compile with g++ and run on fibers: one fiber per CUDA thread, warp collectives
// and barriers as rendezvous points, shared memory as a flat buffer (tests/emu/emu_main.cpp has the runtime).
// Only tests/ builds this; the product never sees it (W2B_EMULATE is defined by tests/emu/Makefile alone).
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n)

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// per-fiber builtins (set by the scheduler before a fiber runs)
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- runtime hooks (emu_main.cpp)
void emu_yield();                                  // give the other fibers a turn
uint64_t emu_warp_exchange(uint64_t v, int op, int arg);  // warp collective: all 32 lanes of the warp call it
void emu_block_barrier(int id, int nthreads);      // __syncthreads (id 0) / bar.sync id, n
enum { EMU_SHFL_XOR = 0, EMU_SHFL_IDX = 1, EMU_BALLOT = 2, EMU_SYNC = 3 };

template <class T>
static inline T emu_shfl(T v, int op, int arg) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  raw = emu_warp_exchange(raw, op, arg);
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
#define EMU_FULL(mask) ((void)(mask))
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o) { EMU_FULL(mask); return emu_shfl(v, EMU_SHFL_XOR, o); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { EMU_FULL(mask); return emu_shfl(v, EMU_SHFL_IDX, src); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { EMU_FULL(mask); return (unsigned)emu_warp_exchange(pred ? 1 : 0, EMU_BALLOT, 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { EMU_FULL(mask); emu_warp_exchange(0, EMU_SYNC, 0); }
static inline void __syncthreads() { emu_block_barrier(0, (int)blockDim.x); }
static inline void __nanosleep(unsigned) { emu_yield(); }
static inline void __threadfence_block() {}

// ---- arithmetic intrinsics: IEEE single operations, no contraction (the file is built with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __frcp_rn(float a) { volatile float r = 1.0f / a; return r; }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
