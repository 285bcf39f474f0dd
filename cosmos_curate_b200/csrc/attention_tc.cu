// softmax(Q K^T / sqrt(d)) V on the tcgen05 tensor cores for the ViT-L/14 shape class: head_dim 64, 129..257 tokens.
//
// 257 = 2 * 128 + 1: the first 256 tokens map onto the tensor cores with no padding at all - two 128-row query tiles
// against one 256-key tile (S = Q K^T is ONE UMMA 128x256x64 per query tile; no online softmax, every key of the row is
// in TMEM at once) - and token 256 is folded in on the SIMT side: its key/value as one extra column per row, its query
// row by a dedicated warp.  Persistent CTAs (one per SM) loop over (image, head) pairs.
//
// Warp roles (352 threads):
//   warp 0      TMA producer: Q tiles (2 x 16 KB), K (32 KB), V (32 KB, double buffered) from the [n*T][3*hidden] QKV matrix
//   warp 1      MMA issuer:   S_g = Q_g K^T  (tcgen05.mma kind::f16, M128 N256 K64) and O_g += P_g[:, chunk] V[chunk]
//                             (M128 N64 K64 per 64-key chunk; V is consumed straight from its [key][dim] rows as an
//                             MN-major B operand, no transpose pass)
//   warps 2-5   softmax group 0 (query rows 0..127), warps 6-9 softmax group 1 (rows 128..255): thread = one row.
//               pass 1: row max over the 256 TMEM columns (+ the extra key), pass 2: exp2 -> fp16 P chunks written to
//               shared memory in the 128-byte-swizzled K-major layout the UMMA A operand wants (2 x 16 KB ring per
//               group, so the exp of chunk c+1 overlaps the MMA of chunk c), epilogue: O / rowsum -> fp16 -> global.
//   warp 10     query row 256 entirely on SIMT (257 dot products of 64 + softmax + 257-term weighted sum from smem K/V).
// TMEM: 512 columns = S_0 | S_1 (256 fp32 columns each); O_g reuses columns 0..63 of S_g once pass 2 has consumed them.
// The two groups run half a phase apart, so one group's exponentials (MUFU) overlap the other group's MMAs.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace cb {

constexpr int kTcThreads = 352;
constexpr int kTileBytes = 128 * 128;       // 128 rows x 64 fp16, SW128
constexpr int kKVBytes = 2 * kTileBytes;    // 256 rows
constexpr int kOffQ = 0;                    // 2 tiles
constexpr int kOffK = 2 * kTileBytes;       // 2 buffers x 256 keys
constexpr int kOffV = kOffK + 2 * kKVBytes; // 2 buffers x 256 keys
constexpr int kOffP = kOffV + 2 * kKVBytes; // [group][2] x 128 rows x 64 keys
constexpr int kOffPx = kOffP + 4 * kTileBytes;
constexpr int kOffX = kOffPx + 1024;         // K and V rows of the extra token: [parity][k|v][64] fp16
constexpr int kOffBar = kOffX + 512;
constexpr int kTcSmem = kOffBar + 256 + 1024 /* alignment slack */;
static_assert(kTcSmem <= 232448, "shared memory budget");

struct AttnTcArgs {
  const __half* qkv;
  __half* out;
  int tokens, heads, n_units;  // n_units = images * heads
  float scale_log2e;
  int debug_skip_max;  // timing experiment only (CB_ATTN_DEBUG_SKIPMAX=1): wrong results
  int use_token;       // CB_ATTN_TOKEN=1: the two softmax groups take strict turns on the exp phase (A/B switch; default off)
  long long* trace;    // CB_ATTN_DEBUG_TRACE=1: clock64 stamps of CTA 0, [group][unit < 16][16 events]
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// volatile: a run of these stays a run of back-to-back MUFU issues (the exp phase is paced by the MUFU pipe, one warp
// instruction per 8 cycles; interleaving the dependent FADDs between them lets a lone warp reach only half that rate)
__device__ __forceinline__ float ex2f_v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// tcgen05.wait::ld that names the destination registers of the load it completes: arithmetic on them cannot be scheduled
// above the wait (a bare wait has no data dependence on them), which makes issuing the NEXT load before the wait safe.
__device__ __forceinline__ void tmem_ld_wait_pin(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                 "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                 "+r"(r[31])::"memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// MN-major B operand, SW128: rows are K (keys), each row = 64 contiguous N elements (128 B); 8-row groups 1024 B apart.
// LBO (stride between 64-element column blocks along N) is irrelevant for N = 64; both strides are set to 1024 B.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// FULL: tokens is 256 or 257, i.e. every column / row of the tensor-core tiles is a real token (no masking code at all)
template <bool FULL>
__global__ void __launch_bounds__(kTcThreads, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_row, const AttnTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  float* px = reinterpret_cast<float*>(smem + kOffPx);
  uint8_t* sX = smem + kOffX;  // [parity][0: k_256 | 1: v_256] 128 B each
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t *q_full = bars, *q_free = bars + 2, *k_full = bars + 4, *k_free = bars + 6, *v_full = bars + 8, *v_free = bars + 10;
  uint64_t *s_ready = bars + 12, *s_free = bars + 14, *o_ready = bars + 16, *p_ready = bars + 18 /*[g*2+b]*/, *p_free = bars + 22;
  uint64_t* tok = bars + 26;  // exp-phase token: the two softmax groups take turns on the MUFU pipe
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.tokens, hidden = a.heads * 64;
  const bool has_extra = T == 257;
  const int t_mma = FULL ? 256 : (T < 256 ? T : 256);  // keys / query rows living in the tensor-core tiles
  const size_t row_stride = (size_t)3 * hidden;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1), mbar_init(&q_free[i], 4), mbar_init(&v_full[i], 1), mbar_init(&v_free[i], 10);
      mbar_init(&k_full[i], 1), mbar_init(&k_free[i], 10);
      mbar_init(&s_ready[i], 1), mbar_init(&s_free[i], 4), mbar_init(&o_ready[i], 1), mbar_init(&tok[i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_ready[i], 4), mbar_init(&p_free[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) tma_prefetch_desc(&map_qkv);
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer
      int it = 0;
      for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, ++it) {
        const int img = u / a.heads, h = u - img * a.heads, row0 = img * T, vb = it & 1;
        // order: what group 0 (half a unit ahead) needs first; Q of group 1 last - its buffer is released late
        auto load_q = [&](int g) {
          mbar_wait_parked(&q_free[g], (it & 1) ^ 1);
          mbar_expect_tx(&q_full[g], kTileBytes);
          tma_load_2d(sQ + g * kTileBytes, &map_qkv, &q_full[g], h * 64, row0 + g * 128);
        };
        load_q(0);
        mbar_wait_parked(&k_free[vb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[vb], kKVBytes + (has_extra ? 128 : 0));
        if (has_extra) tma_load_2d(sX + vb * 256, &map_row, &k_full[vb], hidden + h * 64, row0 + 256);
        tma_load_2d(sK + vb * kKVBytes, &map_qkv, &k_full[vb], hidden + h * 64, row0);
        tma_load_2d(sK + vb * kKVBytes + kTileBytes, &map_qkv, &k_full[vb], hidden + h * 64, row0 + 128);
        mbar_wait_parked(&v_free[vb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[vb], kKVBytes + (has_extra ? 128 : 0));
        if (has_extra) tma_load_2d(sX + vb * 256 + 128, &map_row, &v_full[vb], 2 * hidden + h * 64, row0 + 256);
        tma_load_2d(sV + vb * kKVBytes, &map_qkv, &v_full[vb], 2 * hidden + h * 64, row0);
        tma_load_2d(sV + vb * kKVBytes + kTileBytes, &map_qkv, &v_full[vb], 2 * hidden + h * 64, row0 + 128);
        load_q(1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: an event loop, so that the two softmax groups can run half a unit apart
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 256, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0) | (1u << 16);  // B is MN-major
      const int n_it = a.n_units > (int)blockIdx.x ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      int it_g[2] = {0, 0}, stage[2] = {0, 0};  // stage 0: S = Q K^T pending; 1..4: O += P[chunk] V[chunk] pending
      int qk_cnt[2] = {0, 0}, pv_cnt[2] = {0, 0};  // groups done with the K / V buffer of unit parity
      while (it_g[0] < n_it || it_g[1] < n_it) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int it = it_g[g];
          if (it >= n_it) continue;
          const int kb = it & 1;
          if (stage[g] == 0) {
            if (mbar_test(&q_full[g], it & 1) && mbar_test(&k_full[kb], (it >> 1) & 1) && mbar_test(&s_free[g], (it & 1) ^ 1)) {
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_u32(sQ + g * kTileBytes)), db = umma_desc_sw128(smem_u32(sK + kb * kKVBytes));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_qk, k != 0);
              umma_commit(&s_ready[g]);
              if (++qk_cnt[kb] == 2) qk_cnt[kb] = 0, umma_commit(&k_free[kb]);
              stage[g] = 1, progress = true;
            }
          } else {
            const int c = stage[g] - 1, b = c & 1, use = it * 2 + (c >> 1);
            if ((c != 0 || mbar_test(&v_full[kb], (it >> 1) & 1)) && mbar_test(&p_ready[g * 2 + b], use & 1)) {
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_u32(sP + (g * 2 + b) * kTileBytes));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t db = umma_desc_sw128_mn(smem_u32(sV + kb * kKVBytes + (c * 64 + k * 16) * 128));
                umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db, idesc_pv, (c | k) != 0);
              }
              umma_commit(&p_free[g * 2 + b]);
              if (c == 3) {
                umma_commit(&o_ready[g]);
                if (++pv_cnt[kb] == 2) pv_cnt[kb] = 0, umma_commit(&v_free[kb]);
                stage[g] = 0, it_g[g] = it + 1;
              } else {
                stage[g] = c + 2;
              }
              progress = true;
            }
          }
        }
        (void)progress;  // tight poll: __nanosleep has ~1 us granularity
      }
    }
  } else if (warp < 10) {  // ===== softmax groups: one thread per query row
    const int g = (warp - 2) >> 2, q = warp & 3;  // q = TMEM lane quarter this warp may touch
    const int r = q * 32 + lane, row = g * 128 + r;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 256);
    const uint8_t* q_row = sQ + g * kTileBytes + r * 128;
    int it = 0;
    const bool tracing = a.trace && blockIdx.x == 0 && q == 2 && lane == 0;  // warps 2 and 6
#define CB_TRACE(ev)                                                                  \
  do {                                                                                \
    if (tracing && it < 16) a.trace[(g * 16 + it) * 16 + (ev)] = clock64();           \
  } while (0)
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, ++it) {
      const int img = u / a.heads, h = u - img * a.heads;
      const size_t row0 = (size_t)img * T;
      CB_TRACE(0);
      const int xb = it & 1;
      mbar_wait_parked(&q_full[g], it & 1);  // already complete (the MMA waited on it); taken for the TMA-write -> generic-read ordering
      // score against the extra key (token 256): q_row . k_256, fp32 accumulate (needs Q only, not S)
      float s_x = -INFINITY;
      if (has_extra) {
        mbar_wait_parked(&k_full[xb], (it >> 1) & 1);  // the extra token's K row arrives with the K tile
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 qa = *reinterpret_cast<const uint4*>(q_row + ((j ^ (r & 7)) << 4));
          const uint4 kxj = *reinterpret_cast<const uint4*>(sX + xb * 256 + j * 16);
          const __half2* q2 = reinterpret_cast<const __half2*>(&qa);
          const __half2* k2 = reinterpret_cast<const __half2*>(&kxj);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 qf = __half22float2(q2[e]), kf = __half22float2(k2[e]);
            acc0 = fmaf(qf.x, kf.x, acc0), acc1 = fmaf(qf.y, kf.y, acc1);
          }
        }
        s_x = acc0 + acc1;
      }
      CB_TRACE(1);
      mbar_wait_parked(&s_ready[g], it & 1);
      tc_fence_after();
      CB_TRACE(2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_free[g]), mbar_arrive(&k_free[xb]);  // this warp is done with Q_g and with the extra K row

      // pass 1: row maximum; the load of block cc+1 is in flight while block cc is reduced
      float mx = s_x;
      uint32_t va[32], vb2[32];
      tmem_ld_32x32b_x32(t_row, va);
#pragma unroll 1
      for (int cc = 0; cc < (a.debug_skip_max ? 0 : 8); cc += 2) {
        tmem_ld_wait_pin(va);
        tmem_ld_32x32b_x32(t_row + (uint32_t)((cc + 1) * 32), vb2);
        if (FULL || cc * 32 + 32 <= t_mma) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i < t_mma) mx = fmaxf(mx, __uint_as_float(va[i]));
        }
        tmem_ld_wait_pin(vb2);
        tmem_ld_32x32b_x32(t_row + (uint32_t)(((cc + 2) & 7) * 32), va);  // wraps to block 0: the first block of pass 2
        if (FULL || cc * 32 + 64 <= t_mma) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(vb2[i]), __uint_as_float(vb2[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + 32 + i < t_mma) mx = fmaxf(mx, __uint_as_float(vb2[i]));
        }
      }
      const float mb = mx * a.scale_log2e;
      float sum = has_extra ? ex2f(fmaf(s_x, a.scale_log2e, -mb)) : 0.f;
      const float p_x = sum;
      float sum1 = 0.f;

      // pass 2: P = exp2(S * scale - max) in 64-key chunks -> fp16 -> swizzled shared memory.  `va` already holds (or is
      // receiving) columns 0..31; each half's successor is requested before the exponentials of the current half.
      auto emit = [&](const uint32_t* v, uint8_t* prow, int half, int k0) {
        uint32_t w[16];
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = fmaf(__uint_as_float(v[i]), a.scale_log2e, -mb);
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = ex2f_v(x[i]);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = x[i], p1 = x[i + 1];
          if (!FULL) {
            if (k0 + i >= t_mma) p0 = 0.f;
            if (k0 + i + 1 >= t_mma) p1 = 0.f;
          }
          sum += p0, sum1 += p1;
          w[i >> 1] = pack2(p0, p1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(prow + (((half * 4 + j) ^ (r & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      };
      // optional (CB_ATTN_TOKEN=1): strict alternation of the exp phases of the two groups.  Measured neutral-to-slightly-worse
      // on B200 (0.214 vs 0.207 ms per ViT-L layer): the kernel is bound by issue slots / latency of 2-3 warps per SM
      // sub-partition, not by the MUFU pipe, so it is off by default.
      CB_TRACE(3);
      if (a.use_token) mbar_wait_parked(&tok[g], g == 0 ? (it & 1) ^ 1 : (it & 1));
      CB_TRACE(4);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int b = c & 1, use = it * 2 + (c >> 1);
        uint8_t* prow = sP + (g * 2 + b) * kTileBytes + r * 128;
        mbar_wait_parked(&p_free[g * 2 + b], (use & 1) ^ 1);
        CB_TRACE(5 + 2 * c);
        tmem_ld_wait_pin(va);
        tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + 32), vb2);
        emit(va, prow, 0, c * 64);
        tmem_ld_wait_pin(vb2);
        if (c < 3) tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + 64), va);
        emit(vb2, prow, 1, c * 64 + 32);
        fence_proxy_async();  // generic-proxy writes of P -> visible to the UMMA (async proxy)
        tc_fence_before();    // the TMEM reads above precede the MMA that overwrites columns 0..63 with O
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_ready[g * 2 + b]);
          if (c == 3) mbar_arrive(&tok[g ^ 1]);
        }
        CB_TRACE(6 + 2 * c);
      }
      sum += sum1;

      // epilogue: O / rowsum (+ the extra key's value row) -> fp16 -> global
      mbar_wait_parked(&o_ready[g], it & 1);
      if (has_extra) mbar_wait_parked(&v_full[xb], (it >> 1) & 1);  // long complete; taken for the TMA-write -> generic-read ordering
      tc_fence_after();
      CB_TRACE(13);
      const float inv = 1.0f / sum;
      __half* orow = a.out + (row0 + row) * hidden + h * 64;
      tmem_ld_32x32b_x32(t_row, va);
      tmem_ld_32x32b_x32(t_row + 32u, vb2);
      tmem_ld_wait_pin(va);
      tmem_ld_wait_pin(vb2);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t* v = half ? vb2 : va;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * j + e]);
          if (has_extra) {
            const uint4 vxj = *reinterpret_cast<const uint4*>(sX + xb * 256 + 128 + (half * 4 + j) * 16);
            const __half2* v2 = reinterpret_cast<const __half2*>(&vxj);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 vf = __half22float2(v2[e]);
              o[2 * e] = fmaf(p_x, vf.x, o[2 * e]), o[2 * e + 1] = fmaf(p_x, vf.y, o[2 * e + 1]);
            }
          }
          if (FULL || row < t_mma)
            *reinterpret_cast<uint4*>(orow + half * 32 + j * 8) =
                make_uint4(pack2(o[0] * inv, o[1] * inv), pack2(o[2] * inv, o[3] * inv), pack2(o[4] * inv, o[5] * inv), pack2(o[6] * inv, o[7] * inv));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]), mbar_arrive(&v_free[xb]);  // done with TMEM and with the extra V row
      CB_TRACE(14);
    }
  } else {  // ===== warp 10: query row 256 on SIMT
    int it = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, ++it) {
      const int img = u / a.heads, h = u - img * a.heads, vb = it & 1;
      const size_t row0 = (size_t)img * T;
      mbar_wait_parked(&k_full[vb], (it >> 1) & 1);
      float s[8], s_x = 0.f;
      if (has_extra) {
        const __half* xrow = a.qkv + (row0 + 256) * row_stride + h * 64;
        uint4 qx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) qx[j] = __ldg(reinterpret_cast<const uint4*>(xrow) + j);
        auto dot = [&](const uint4* krow, bool swz, int key) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 kb = swz ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(krow) + ((j ^ (key & 7)) << 4)) : __ldg(krow + j);
            const __half2* q2 = reinterpret_cast<const __half2*>(&qx[j]);
            const __half2* k2 = reinterpret_cast<const __half2*>(&kb);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 qf = __half22float2(q2[e]), kf = __half22float2(k2[e]);
              acc = fmaf(qf.x, kf.x, acc), acc = fmaf(qf.y, kf.y, acc);
            }
          }
          return acc;
        };
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = lane + 32 * i;
          s[i] = dot(reinterpret_cast<const uint4*>(sK + vb * kKVBytes + key * 128), true, key);
        }
        s_x = dot(reinterpret_cast<const uint4*>(xrow + hidden), false, 0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_free[vb]);
      if (has_extra) {
        float mx = s_x;
#pragma unroll
        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, s[i]);
        for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float mb = mx * a.scale_log2e;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p = ex2f(fmaf(s[i], a.scale_log2e, -mb));
          px[lane + 32 * i] = p;
          sum += p;
        }
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float p_x = ex2f(fmaf(s_x, a.scale_log2e, -mb));
        sum += p_x;
        __syncwarp();
        mbar_wait_parked(&v_full[vb], (it >> 1) & 1);
        // lane owns output dims 2*lane, 2*lane+1: byte lane*4 of every V row -> 16-byte chunk lane>>2, offset (lane&3)*4
        const uint8_t* vbase = sV + vb * kKVBytes + (lane & 3) * 4;
        const int ch = lane >> 2;
        float o0 = 0.f, o1 = 0.f, oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int key = 0; key < 256; key += 4) {
          const float4 p4 = *reinterpret_cast<const float4*>(px + key);
          const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // four independent accumulator pairs: the loop is not an FMA latency chain
            const int kk = key + e;
            const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vbase + kk * 128 + ((ch ^ (kk & 7)) << 4)));
            oa[e] = fmaf(pv[e], vf.x, oa[e]), ob[e] = fmaf(pv[e], vf.y, ob[e]);
          }
        }
        o0 = (oa[0] + oa[1]) + (oa[2] + oa[3]), o1 = (ob[0] + ob[1]) + (ob[2] + ob[3]);
        const float2 vxf = __half22float2(*reinterpret_cast<const __half2*>(a.qkv + (row0 + 256) * row_stride + 2 * hidden + h * 64 + 2 * lane));
        o0 = fmaf(p_x, vxf.x, o0), o1 = fmaf(p_x, vxf.y, o1);
        const float inv = 1.0f / sum;
        *reinterpret_cast<uint32_t*>(a.out + (row0 + 256) * hidden + h * 64 + 2 * lane) = pack2(o0 * inv, o1 * inv);
      } else {
        mbar_wait_parked(&v_full[vb], (it >> 1) & 1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_free[vb]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host side: returns CB_OK after launching, or -1 ("not my shape") so the caller falls back to the mma.sync kernel
int attention_tc(cb_ctx* ctx, const void* qkv, void* out, int n, int tokens, int heads, int head_dim, cudaStream_t stream, bool* launched) {
  *launched = false;
  const char* sel = std::getenv("CB_ATTN_KERNEL");
  if (sel && std::strcmp(sel, "mma") == 0) return CB_OK;
  if (head_dim != 64 || tokens < 129 || tokens > 257) return CB_OK;
  const int hidden = heads * 64;
  CUtensorMap map;
  const uint64_t dims[2] = {(uint64_t)3 * hidden, (uint64_t)n * tokens}, strides[1] = {(uint64_t)3 * hidden * 2};
  const uint32_t box[2] = {64, 128};
  int rc = make_tensor_map(ctx, &map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  CUtensorMap map_row;  // one token's 64-element slice (the extra token's K and V rows), unswizzled
  const uint32_t box_row[2] = {64, 1};
  rc = make_tensor_map(ctx, &map_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qkv, dims, strides, box_row, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  static bool attr_done[64] = {};  // the attribute is per device: one process may drive several
  bool& attr_set = attr_done[ctx->device & 63];
  if (!attr_set) {
    CB_CUDA(ctx, cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    CB_CUDA(ctx, cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    attr_set = true;
  }
  const char* dbg = std::getenv("CB_ATTN_DEBUG_SKIPMAX");
  AttnTcArgs a{(const __half*)qkv, (__half*)out, tokens, heads, n * heads, 1.4426950408889634f / sqrtf(64.f), dbg && dbg[0] == '1', 0, nullptr};
  const char* tk = std::getenv("CB_ATTN_TOKEN");
  a.use_token = tk && tk[0] == '1';
  const char* trc = std::getenv("CB_ATTN_DEBUG_TRACE");
  const bool tracing = trc && trc[0] == '1';
  if (tracing) {
    CB_CUDA(ctx, cudaMalloc(&a.trace, 2 * 16 * 16 * sizeof(long long)));
    CB_CUDA(ctx, cudaMemset(a.trace, 0, 2 * 16 * 16 * sizeof(long long)));
  }
  const int grid = std::min(n * heads, ctx->sm_count);
  mark_launch(ctx, CB_PROF_ATTENTION, stream);
  if (tokens >= 256)
    attention_tc_kernel<true><<<grid, kTcThreads, kTcSmem, stream>>>(map, map_row, a);
  else
    attention_tc_kernel<false><<<grid, kTcThreads, kTcSmem, stream>>>(map, map_row, a);
  CB_CUDA(ctx, cudaGetLastError());
  if (tracing) {
    long long h[2 * 16 * 16];
    CB_CUDA(ctx, cudaStreamSynchronize(stream));
    CB_CUDA(ctx, cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(a.trace);
    const long long t0 = h[0];
    for (int g = 0; g < 2; ++g)
      for (int it = 0; it < 8; ++it) {
        printf("trace g%d it%d:", g, it);
        for (int e = 0; e < 15; ++e) printf(" %lld", h[(g * 16 + it) * 16 + e] ? h[(g * 16 + it) * 16 + e] - t0 : -1);
        printf("\n");
      }
  }
  *launched = true;
  return CB_OK;
}

}  // namespace cb
