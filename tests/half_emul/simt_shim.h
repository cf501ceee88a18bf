// simt_shim.h — TEST INFRASTRUCTURE ONLY.
//
// Lets alpharat_b200/csrc/mcts_half.cuh (the two-trees-per-warp device code) compile with g++ and run on
// the host: the 16 lanes of a half are 16 cooperative fibers (ucontext), every warp collective
// (__shfl*_sync, __ballot_sync, __syncwarp) is a rendezvous of the whole half.  Each lane publishes its
// operand and yields to the next lane; when it is resumed every lane of the half has published.  The
// rendezvous also CHECKS the invariant the kernel's __activemask() member masks rest on: every lane of a
// half executes the same sequence of collectives (same kind, same order) — a lane that arrives at a
// different collective, or finishes early, aborts the run.
//
// Lanes run one after the other between two collectives, not in lockstep.  Code that is only correct because a warp
// issues a load for all lanes before a later store of one lane (say, every lane reading a child-table word that lane 0
// then overwrites, with no collective in between) therefore fails here although it works on the GPU: such a pattern is a
// formal race under independent thread scheduling, and an experiment that introduced one was caught this way.
// -DSIMT_BACKTRACE (with -O0 -g) prints the failing lane's return addresses and every lane's collective count.
//
// The product never includes this file (alpharat_b200/ has no reference to tests/).
#pragma once
#define AR_HOST_EMUL 1

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <algorithm>
#ifdef SIMT_BACKTRACE
#include <execinfo.h>
#endif

// ---- CUDA keywords -------------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __builtin_assume(x) ((void)0)
#define __isShared(p) true
#define __isGlobal(p) true

struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- scalar intrinsics ---------------------------------------------------------------------------------
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline int __float2int_rn(float f) { return (int)lrintf(f); }
static inline uint32_t __float2uint_rz(float f) {  // cvt.rzi.u32.f32: saturating, NaN -> 0
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xffffffffu;
  return (uint32_t)f;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { unsigned int o = *p; *p += v; return o; }
static inline int atomicCAS(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
using std::max;
using std::min;

// ---- the half: 16 fibers and their rendezvous -----------------------------------------------------------
namespace simt {

constexpr int MAX_LANES = 32;
enum Kind : uint32_t { K_SHFL = 1, K_SHFL_XOR, K_SHFL_UP, K_SHFL_DOWN, K_BALLOT, K_SYNC, K_REDUCE, K_VOTE };

// Fibre switch.  x86-64: save the callee-saved registers and swap stack pointers (swapcontext costs a system call per
// switch — it saves the signal mask — and a search makes tens of thousands of them); elsewhere: ucontext.
#if defined(__x86_64__)
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".hidden simt_switch\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch,.-simt_switch\n");
typedef void* fiber_t;
#else
typedef ucontext_t fiber_t;
#endif

struct Half {  // a group of lanes that executes collectives together: a half of a warp (16) or a whole warp (32)
  fiber_t ctx[MAX_LANES];
  fiber_t main_ctx;
  void* stacks[MAX_LANES] = {nullptr};
  int step = 1;           // rotation order of the lanes: +1, or -1 (SIMT_REVERSE=1: a result that depends on it is a race)
  int lanes = 16;         // 16: a half (mcts_half.cuh); 32: a warp (the one-tree-per-warp code of mcts_device.cuh)
  int cur = 0;            // lane that is running
  int hbase = 0;          // 0 or 16: which half of the warp these lanes are (0 for a whole warp)
  uint64_t seq[MAX_LANES] = {0};
  uint32_t slot[2][MAX_LANES];
  uint32_t kind[2][MAX_LANES];
  bool done[MAX_LANES] = {false};
  int n_done = 0;
  void (*body)(int lane, void* arg) = nullptr;
  void* arg = nullptr;
  uint64_t collectives = 0;
};
static Half* g_half = nullptr;

static inline void switch_to(fiber_t* from, fiber_t* to) {
#if defined(__x86_64__)
  simt_switch(from, *to);
#else
  swapcontext(from, to);
#endif
}

static inline void fail(const char* what) {
  fprintf(stderr, "simt emulation: %s (lane %d)\n", what, g_half ? g_half->cur : -1);
#ifdef SIMT_BACKTRACE  // g++ -O0 -g -DSIMT_BACKTRACE: return addresses of the failing lane for addr2line
  void* bt[48];
  const int n = backtrace(bt, 48);
  backtrace_symbols_fd(bt, n, 2);
  if (g_half)
    for (int i = 0; i < g_half->lanes; ++i)
      fprintf(stderr, "  lane %2d: collectives %llu, last kinds %u %u\n", i, (unsigned long long)g_half->seq[i],
              g_half->kind[0][i], g_half->kind[1][i]);
#endif
  abort();
}

// Publish one operand of collective `k` and let every other lane reach the same point.
static inline const uint32_t* rendezvous(uint32_t k, uint32_t v) {
  Half& h = *g_half;
  const int me = h.cur;
  const uint64_t s = h.seq[me]++;
  const int par = (int)(s & 1);
  h.slot[par][me] = v;
  h.kind[par][me] = k;
  if (me == 0) h.collectives += 1;
  const int LANES = h.lanes;
  const int next = (me + h.step + LANES) % LANES;
  if (h.done[next]) fail("a lane finished while another is still executing collectives");
  h.cur = next;
  switch_to(&h.ctx[me], &h.ctx[next]);
  // resumed: the rotation has come back, every lane has published collective s
  for (int i = 0; i < LANES; ++i) {
    if (h.seq[i] <= s) fail("a lane did not reach this collective: control flow diverged inside a half");
    if (h.kind[par][i] != k) fail("lanes of a half are at different collectives");
  }
  return h.slot[par];
}

static void trampoline(int lane) {
  Half& h = *g_half;
  const int LANES = h.lanes;
  h.body(lane, h.arg);
  h.done[lane] = true;
  h.n_done += 1;
  // every lane finishes after the same number of collectives; the last one returns to the caller
  for (int i = 0; i < LANES; ++i)
    if (!h.done[i] && h.seq[i] != h.seq[lane]) fail("lanes finished after different numbers of collectives");
  fiber_t dead;
  if (h.n_done == LANES) {
    h.cur = -1;
    switch_to(&dead, &h.main_ctx);
  }
  const int next = (lane + h.step + LANES) % LANES;
  h.cur = next;
  switch_to(&dead, &h.ctx[next]);
  abort();  // a finished lane is never resumed
}
#if defined(__x86_64__)
static void fiber_entry() { trampoline(g_half->cur); }
#endif

// Run body(lane, arg) on the 16 lanes of a half (hbase 0 or 16), or on the 32 lanes of a warp, to completion.
static inline uint64_t run_half(int hbase, void (*body)(int, void*), void* arg, int lanes = 16) {
  static Half h;
  g_half = &h;
  const int LANES = lanes;
  h.lanes = lanes;
  h.step = getenv("SIMT_REVERSE") ? -1 : 1;
  h.hbase = hbase;
  h.body = body;
  h.arg = arg;
  h.n_done = 0;
  h.collectives = 0;
  const size_t STACK = 1 << 20;
  for (int i = 0; i < LANES; ++i) {
    h.seq[i] = 0;
    h.done[i] = false;
    if (!h.stacks[i]) h.stacks[i] = malloc(STACK);
#if defined(__x86_64__)
    // initial frame: six callee-saved registers, then the entry point as the return address; the stack pointer is
    // 8 mod 16 when fiber_entry starts, as after a call
    uintptr_t top = ((uintptr_t)h.stacks[i] + STACK) & ~(uintptr_t)15;
    void** sp = reinterpret_cast<void**>(top - 16);
    *sp = reinterpret_cast<void*>(&fiber_entry);
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    h.ctx[i] = sp;
#else
    getcontext(&h.ctx[i]);
    h.ctx[i].uc_stack.ss_sp = h.stacks[i];
    h.ctx[i].uc_stack.ss_size = STACK;
    h.ctx[i].uc_link = nullptr;
    makecontext(&h.ctx[i], (void (*)())trampoline, 1, i);
#endif
  }
  h.cur = h.step > 0 ? 0 : LANES - 1;
  switch_to(&h.main_ctx, &h.ctx[h.cur]);
  return h.collectives;
}

static inline int lane_in_half() { return g_half->cur; }

}  // namespace simt

// ---- warp collectives (member mask: the half that is executing, as in the kernel) -------------------------
static inline unsigned group_mask() { return simt::g_half->lanes == 32 ? 0xffffffffu : (0xffffu << simt::g_half->hbase); }
static inline unsigned __activemask() { return group_mask(); }
static inline void check_member_mask(unsigned mask) {
  if (mask != group_mask()) simt::fail("collective with a member mask that is not the executing half / warp");
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  check_member_mask(mask);
  simt::rendezvous(simt::K_SYNC, 0);
}
static inline uint32_t __shfl_sync(unsigned mask, uint32_t v, int src, int width = 32) {
  check_member_mask(mask);
  if (width > simt::g_half->lanes) simt::fail("shuffle wider than the emulated group");
  const int me = simt::lane_in_half();
  const uint32_t* s = simt::rendezvous(simt::K_SHFL, v);
  return s[(me & ~(width - 1)) | (src & (width - 1))];
}
static inline int __shfl_sync(unsigned mask, int v, int src, int width = 32) { return (int)__shfl_sync(mask, (uint32_t)v, src, width); }
static inline float __shfl_sync(unsigned mask, float v, int src, int width = 32) {
  return __uint_as_float(__shfl_sync(mask, __float_as_uint(v), src, width));
}
static inline uint32_t __shfl_xor_sync(unsigned mask, uint32_t v, int lanemask, int width = 32) {
  check_member_mask(mask);
  if (width > simt::g_half->lanes) simt::fail("shuffle wider than the emulated group");
  const int me = simt::lane_in_half();
  const uint32_t* s = simt::rendezvous(simt::K_SHFL_XOR, v);
  const int src = me ^ lanemask;
  return ((src & ~(width - 1)) == (me & ~(width - 1))) ? s[src] : s[me];
}
static inline float __shfl_xor_sync(unsigned mask, float v, int lanemask, int width = 32) {
  return __uint_as_float(__shfl_xor_sync(mask, __float_as_uint(v), lanemask, width));
}
static inline uint32_t __shfl_up_sync(unsigned mask, uint32_t v, unsigned delta, int width = 32) {
  check_member_mask(mask);
  if (width > simt::g_half->lanes) simt::fail("shuffle wider than the emulated group");
  const int me = simt::lane_in_half();
  const uint32_t* s = simt::rendezvous(simt::K_SHFL_UP, v);
  return ((me & (width - 1)) >= (int)delta) ? s[me - (int)delta] : s[me];
}
static inline float __shfl_up_sync(unsigned mask, float v, unsigned delta, int width = 32) {
  return __uint_as_float(__shfl_up_sync(mask, __float_as_uint(v), delta, width));
}
static inline uint32_t __shfl_down_sync(unsigned mask, uint32_t v, unsigned delta, int width = 32) {
  check_member_mask(mask);
  if (width > simt::g_half->lanes) simt::fail("shuffle wider than the emulated group");
  const int me = simt::lane_in_half();
  const uint32_t* s = simt::rendezvous(simt::K_SHFL_DOWN, v);
  return ((me & (width - 1)) + (int)delta < width) ? s[me + (int)delta] : s[me];
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  check_member_mask(mask);
  const uint32_t* s = simt::rendezvous(simt::K_BALLOT, pred ? 1u : 0u);
  unsigned b = 0;
  for (int i = 0; i < simt::g_half->lanes; ++i) b |= (s[i] & 1u) << i;
  return b << simt::g_half->hbase;
}
// used by the one-tree-per-warp code of mcts_device.cuh (32-lane groups)
static inline int __any_sync(unsigned mask, int pred) {
  check_member_mask(mask);
  const uint32_t* s = simt::rendezvous(simt::K_VOTE, pred ? 1u : 0u);
  int r = 0;
  for (int i = 0; i < simt::g_half->lanes; ++i) r |= (int)(s[i] & 1u);
  return r;
}
static inline int __all_sync(unsigned mask, int pred) {
  check_member_mask(mask);
  const uint32_t* s = simt::rendezvous(simt::K_VOTE, pred ? 1u : 0u);
  int r = 1;
  for (int i = 0; i < simt::g_half->lanes; ++i) r &= (int)(s[i] & 1u);
  return r;
}
static inline uint32_t __reduce_max_sync(unsigned mask, uint32_t v) {
  check_member_mask(mask);
  const uint32_t* s = simt::rendezvous(simt::K_REDUCE, v);
  uint32_t m = 0;
  for (int i = 0; i < simt::g_half->lanes; ++i) m = s[i] > m ? s[i] : m;
  return m;
}
static inline uint32_t __reduce_min_sync(unsigned, uint32_t) { simt::fail("__reduce_min_sync in the half emulation"); return 0; }
static inline uint32_t __reduce_add_sync(unsigned, uint32_t) { simt::fail("__reduce_add_sync in the half emulation"); return 0; }
static inline uint32_t __reduce_or_sync(unsigned, uint32_t) { simt::fail("__reduce_or_sync in the half emulation"); return 0; }
static inline unsigned __match_any_sync(unsigned, uint32_t) { simt::fail("__match_any_sync in the half emulation"); return 0; }
