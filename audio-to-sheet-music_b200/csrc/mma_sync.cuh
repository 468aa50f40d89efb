// Warp-level tensor-core helpers (mma.sync.m16n8k16 bf16 -> fp32) for the small, bandwidth-bound contractions of the
// path (K = 16..1152 on 16-row tiles that already sit in shared memory / registers).  The large compute-bound GEMMs
// use tcgen05 (gemm_tc.cu); a TMEM round trip per 16..128-row tile of these layers would cost more than the math.
#pragma once
#include "common.cuh"

namespace athtd {

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t ld_b32(const bf16* p) { return *(const uint32_t*)p; }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *(uint32_t*)&t;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) { return __bfloat1622float2(*(const __nv_bfloat162*)&v); }

// A fragment (16 rows x 16 k) of a row-major bf16 matrix (pitch P elements); lane = 4*g + q owns rows g, g+8 and
// k columns 2q, 2q+1, 2q+8, 2q+9.  Works on shared or global pointers.
__device__ __forceinline__ void frag_a(const bf16* S, int P, int r0, int k0, int lane, uint32_t (&a)[4]) {
  const int g = lane >> 2, q = lane & 3;
  a[0] = ld_b32(S + (r0 + g) * P + k0 + 2 * q);
  a[1] = ld_b32(S + (r0 + g + 8) * P + k0 + 2 * q);
  a[2] = ld_b32(S + (r0 + g) * P + k0 + 2 * q + 8);
  a[3] = ld_b32(S + (r0 + g + 8) * P + k0 + 2 * q + 8);
}
// A fragment through ldmatrix.x4 (shared memory only; rows 16-byte aligned)
__device__ __forceinline__ void ldsm_a(const bf16* S, int P, int r0, int k0, int lane, uint32_t (&a)[4]) {
  const bf16* p = S + (r0 + (lane & 15)) * P + k0 + ((lane >> 4) << 3);
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
// 16 x 16 block of a row-major tile (rows r0.., columns c0..) through ldmatrix.x4.trans: t[0] = (rows 0-7, cols 0-7)^T, t[1] = (rows 0-7,
// cols 8-15)^T, t[2] = (rows 8-15, cols 0-7)^T, t[3] = (rows 8-15, cols 8-15)^T.  For X = the block: {t0,t1,t2,t3} is the A fragment of
// X^T (m = column, k = row) and {t0,t2} / {t1,t3} are the B fragments (k = row) of columns 0-7 / 8-15: X^T X needs no other load.
__device__ __forceinline__ void ldsm_x4_trans(const bf16* S, int P, int r0, int c0, int lane, uint32_t (&t)[4]) {
  const int m = lane >> 3, r = lane & 7;
  const bf16* p = S + (r0 + ((m >> 1) << 3) + r) * P + c0 + ((m & 1) << 3);
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]) : "r"(addr));
}
// B fragment (16 k x 8 n) from weights stored [n][k] (k contiguous, pitch KP): lane owns n = g, k = 2q, 2q+1, 2q+8, 2q+9
__device__ __forceinline__ void frag_b(const bf16* W, int KP, int n0, int k0, int lane, uint32_t (&b)[2]) {
  const int g = lane >> 2, q = lane & 3;
  b[0] = ld_b32(W + (n0 + g) * KP + k0 + 2 * q);
  b[1] = ld_b32(W + (n0 + g) * KP + k0 + 2 * q + 8);
}
// two B fragments (k0 and k0+16) through one ldmatrix.x4 (shared memory only)
__device__ __forceinline__ void ldsm_b2(const bf16* W, int KP, int n0, int k0, int lane, uint32_t (&b0)[2], uint32_t (&b1)[2]) {
  const bf16* p = W + (n0 + (lane & 7)) * KP + k0 + ((lane >> 3) << 3);
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(b0[0]), "=r"(b0[1]), "=r"(b1[0]), "=r"(b1[1]) : "r"(addr));
}
// B fragments at the same k0 of TWO n-tiles (rows nA.., nB.. of W) through one ldmatrix.x4 (shared memory only)
__device__ __forceinline__ void ldsm_b_pair(const bf16* W, int KP, int nA, int nB, int k0, int lane, uint32_t (&bA)[2], uint32_t (&bB)[2]) {
  const int m = lane >> 3;
  const bf16* p = W + (((m >> 1) ? nB : nA) + (lane & 7)) * KP + k0 + ((m & 1) << 3);
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(bA[0]), "=r"(bA[1]), "=r"(bB[0]), "=r"(bB[1]) : "r"(addr));
}

__device__ __forceinline__ float sigmoid_fast_(float x) { return __frcp_rn(1.0f + __expf(-x)); }

}  // namespace athtd
