// Second-generation tcgen05 attention for head_dim 64 / 129..257 tokens: same tiling as attention_tc.cu (two 128-row query
// tiles x one 256-key tile per (image, head), S in TMEM, V as MN-major B operand, token 256 folded in on SIMT), but TWO
// threads per query row.  The first kernel gave each row to one thread; with 8 softmax warps per SM (2 per scheduler) it was
// bound by issue slots and latency (clock64 trace in DESIGN.md), not by MUFU, TMEM or the tensor pipe.  Here 16 softmax warps
// (4 per scheduler) each own 32 rows x 128 columns: the row maximum and the row sum are combined across the two half-rows
// through shared memory (one named barrier per group each), every half-row exponentiates its own two 64-key chunks into its
// own P buffer, and the epilogue writes 32 of the 64 output dims per thread.
//
// Warp roles (704 threads): 0 TMA producer, 1 MMA issuer (event loop), 2, 3, 20, 21 query row 256 on SIMT (64 keys each),
// 4..19 softmax: group g = (w-4)/8 (query tile), half hf = ((w-4)/4)%2 (key columns hf*128..+127), TMEM lane quarter w%4.
// The extra row used to be ONE warp (257 dot products + a 257-term weighted sum, ~8k cycles per unit): every K and V
// buffer release waited for it, and it - not the exponentials - set the pace of the whole kernel.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace cb {
namespace tc2 {

constexpr int kThreads = 704;  // 22 warps
constexpr int kTileBytes = 128 * 128;     // 128 rows x 64 fp16, SW128
constexpr int kKVBytes = 2 * kTileBytes;  // 256 rows
constexpr int kOffQ = 0;                  // 2 tiles
constexpr int kOffK = 2 * kTileBytes;
constexpr int kOffV = kOffK + kKVBytes;      // 2 buffers
constexpr int kOffP = kOffV + 2 * kKVBytes;  // [group][half] x 128 rows x 64 keys
constexpr int kOffPx = kOffP + 4 * kTileBytes;
constexpr int kOffX = kOffPx + 1024;      // K, V and Q rows of the extra token: [parity][k|v|q][64] fp16 (384 B per parity)
constexpr int kOffExch = kOffX + 1024;    // [max|sum|sx][group][half][128] fp32
constexpr int kOffXs = kOffExch + 3 * 2 * 2 * 128 * 4;  // scratch of the extra-row warps: max[4] sum[4] sx, then o[4][64]
constexpr int kOffBar = kOffXs + 2048;
constexpr int kSmem = kOffBar + 256 + 1024 /* alignment slack */;
static_assert(kSmem <= 232448, "shared memory budget");

struct Args {
  const __half* qkv;
  __half* out;
  int tokens, heads, n_units;
  float scale_log2e;
  long long* trace;  // CB_ATTN_DEBUG_TRACE=1: clock64 stamps of CTA 0: [3 roles][16 units][16 events]
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_ld_wait_pin(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                 "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                 "+r"(r[31])::"memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {  // see attention_tc.cu
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

template <bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
    attention_tc2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_row, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  float* px = reinterpret_cast<float*>(smem + kOffPx);
  uint8_t* sX = smem + kOffX;
  float* ex_max = reinterpret_cast<float*>(smem + kOffExch);  // [g][hf][128]
  float* ex_sum = ex_max + 512;
  float* ex_sx = ex_sum + 512;  // [g][128] (only the first 256 floats used)
  float* xs = reinterpret_cast<float*>(smem + kOffXs);  // [0..3] max, [4..7] sum, [8] s_x
  float* xo = xs + 16;                                  // [4][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t *q_full = bars, *q_free = bars + 2, *k_full = bars + 4, *k_free = bars + 5, *v_full = bars + 6, *v_free = bars + 8;
  uint64_t *s_ready = bars + 10, *s_free = bars + 12, *o_ready = bars + 14, *p_ready = bars + 16 /*[g*2+hf]*/, *p_free = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.tokens, hidden = a.heads * 64;
  const bool has_extra = T == 257;
  const int t_mma = FULL ? 256 : (T < 256 ? T : 256);
  const size_t row_stride = (size_t)3 * hidden;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1), mbar_init(&q_free[i], 4), mbar_init(&v_full[i], 1), mbar_init(&v_free[i], 21);
      mbar_init(&s_ready[i], 1), mbar_init(&s_free[i], 8), mbar_init(&o_ready[i], 1);
    }
    mbar_init(k_full, 1), mbar_init(k_free, 13);
    for (int i = 0; i < 4; ++i) mbar_init(&p_ready[i], 4), mbar_init(&p_free[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) tma_prefetch_desc(&map_qkv), tma_prefetch_desc(&map_row);
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
        auto load_q = [&](int g) {
          mbar_wait_parked(&q_free[g], (it & 1) ^ 1);
          mbar_expect_tx(&q_full[g], kTileBytes);
          tma_load_2d(sQ + g * kTileBytes, &map_qkv, &q_full[g], h * 64, row0 + g * 128);
        };
        load_q(0);
        mbar_wait_parked(k_free, (it & 1) ^ 1);
        mbar_expect_tx(k_full, kKVBytes + (has_extra ? 256 : 0));
        if (has_extra) {
          tma_load_2d(sX + vb * 384, &map_row, k_full, hidden + h * 64, row0 + 256);
          tma_load_2d(sX + vb * 384 + 256, &map_row, k_full, h * 64, row0 + 256);
        }
        tma_load_2d(sK, &map_qkv, k_full, hidden + h * 64, row0);
        tma_load_2d(sK + kTileBytes, &map_qkv, k_full, hidden + h * 64, row0 + 128);
        load_q(1);
        mbar_wait_parked(&v_free[vb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[vb], kKVBytes + (has_extra ? 128 : 0));
        if (has_extra) tma_load_2d(sX + vb * 384 + 128, &map_row, &v_full[vb], 2 * hidden + h * 64, row0 + 256);
        tma_load_2d(sV + vb * kKVBytes, &map_qkv, &v_full[vb], 2 * hidden + h * 64, row0);
        tma_load_2d(sV + vb * kKVBytes + kTileBytes, &map_qkv, &v_full[vb], 2 * hidden + h * 64, row0 + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer (event loop over 2 groups x 2 half-row streams)
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 256, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0) | (1u << 16);  // B is MN-major
      const int n_it = a.n_units > (int)blockIdx.x ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      int it_g[2] = {0, 0};
      int qk_done[2] = {0, 0};   // S issued for the current unit of group g
      int nxt[2][2] = {{0, 0}, {0, 0}};  // next chunk (0, 1, 2 = done) of stream [g][hf]
      int v_ok[2] = {0, 0};
      int qk_cnt = 0, pv_cnt[2] = {0, 0};
      while (it_g[0] < n_it || it_g[1] < n_it) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int it = it_g[g];
          if (it >= n_it) continue;
          const int vb = it & 1;
          if (!qk_done[g]) {
            if (mbar_test(&q_full[g], it & 1) && mbar_test(k_full, it & 1) && mbar_test(&s_free[g], (it & 1) ^ 1)) {
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_u32(sQ + g * kTileBytes)), db = umma_desc_sw128(smem_u32(sK));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_qk, k != 0);
              umma_commit(&s_ready[g]);
              if (++qk_cnt == 2) qk_cnt = 0, umma_commit(k_free);
              qk_done[g] = 1, v_ok[g] = 0, progress = true;
            }
            continue;
          }
          if (!v_ok[g]) {
            if (!mbar_test(&v_full[vb], (it >> 1) & 1)) continue;
            v_ok[g] = 1;
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int i = nxt[g][hf];
            if (i >= 2) continue;
            const int c = hf * 2 + i, use = it * 2 + i;
            // Fixed issue order 0, 2, 1, 3 (chunk i of the second half-row stream right after chunk i of the first): the fp32
            // accumulation order of O - and with it every output bit - is the same on every run.  (O aliases the S columns of
            // chunk 0, so chunk 0 had to be first anyway.)
            if (2 * i + hf != nxt[g][0] + nxt[g][1]) continue;
            if (!mbar_test(&p_ready[g * 2 + hf], use & 1)) continue;
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(smem_u32(sP + (g * 2 + hf) * kTileBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t db = umma_desc_sw128_mn(smem_u32(sV + vb * kKVBytes + (c * 64 + k * 16) * 128));
              umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db, idesc_pv, (c | k) != 0);
            }
            umma_commit(&p_free[g * 2 + hf]);
            nxt[g][hf] = i + 1, progress = true;
            if (nxt[g][0] == 2 && nxt[g][1] == 2) {
              umma_commit(&o_ready[g]);
              if (++pv_cnt[vb] == 2) pv_cnt[vb] = 0, umma_commit(&v_free[vb]);
              nxt[g][0] = nxt[g][1] = 0, qk_done[g] = 0, it_g[g] = it + 1;
            }
          }
        }
        (void)progress;  // tight poll: __nanosleep has ~1 us granularity, which is a quarter of a unit
      }
    }
  } else if (warp < 4 || warp >= 20) {  // ===== query row 256 on SIMT: four warps, 64 keys each
    const int xw = warp < 4 ? warp - 2 : warp - 18;
    auto xsync = [] { asm volatile("bar.sync 3, 128;" ::: "memory"); };
    int it = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, ++it) {
      const int img = u / a.heads, h = u - img * a.heads, vb = it & 1;
      const size_t row0 = (size_t)img * T;
      const bool xtr = a.trace && blockIdx.x == 0 && xw == 0 && lane == 0 && it < 16;
      if (xtr) a.trace[(2 * 16 + it) * 16 + 0] = clock64();
      mbar_wait_parked(k_full, it & 1);
      if (xtr) a.trace[(2 * 16 + it) * 16 + 1] = clock64();
      float s0 = 0.f, s1 = 0.f, s_x = 0.f;
      const int j0 = xw * 64 + lane, j1 = j0 + 32;
      if (has_extra) {
        uint4 qx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) qx[j] = *reinterpret_cast<const uint4*>(sX + vb * 384 + 256 + j * 16);
        auto dot = [&](const uint8_t* krow, int swz) {
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 kb = *reinterpret_cast<const uint4*>(krow + ((j ^ swz) << 4));
            const __half2* q2 = reinterpret_cast<const __half2*>(&qx[j]);
            const __half2* k2 = reinterpret_cast<const __half2*>(&kb);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 qf = __half22float2(q2[e]), kf = __half22float2(k2[e]);
              acc0 = fmaf(qf.x, kf.x, acc0), acc1 = fmaf(qf.y, kf.y, acc1);
            }
          }
          return acc0 + acc1;
        };
        s0 = dot(sK + j0 * 128, j0 & 7);
        s1 = dot(sK + j1 * 128, j1 & 7);
        float m = fmaxf(s0, s1);
        if (xw == 0) s_x = dot(sX + vb * 384, 0), m = fmaxf(m, s_x);
        for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) {
          xs[xw] = m;
          if (xw == 0) xs[8] = s_x;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(k_free);
      float mb = 0.f;
      if (has_extra) {
        xsync();
        mb = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3])) * a.scale_log2e;
        const float p0 = ex2f(fmaf(s0, a.scale_log2e, -mb)), p1 = ex2f(fmaf(s1, a.scale_log2e, -mb));
        px[j0] = p0, px[j1] = p1;
        float sum = p0 + p1;
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) xs[4 + xw] = sum;
        __syncwarp();
      }
      if (xtr) a.trace[(2 * 16 + it) * 16 + 2] = clock64();
      mbar_wait_parked(&v_full[vb], (it >> 1) & 1);
      if (xtr) a.trace[(2 * 16 + it) * 16 + 3] = clock64();
      if (has_extra) {
        // this warp's 64 keys; lane owns output dims 2*lane, 2*lane+1: byte lane*4 of every V row
        const uint8_t* vbase = sV + vb * kKVBytes + (lane & 3) * 4;
        const int ch = lane >> 2;
        float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int key = xw * 64; key < xw * 64 + 64; key += 4) {
          const float4 p4 = *reinterpret_cast<const float4*>(px + key);
          const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kk = key + e;
            const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vbase + kk * 128 + ((ch ^ (kk & 7)) << 4)));
            oa[e] = fmaf(pv[e], vf.x, oa[e]), ob[e] = fmaf(pv[e], vf.y, ob[e]);
          }
        }
        xo[xw * 64 + 2 * lane] = (oa[0] + oa[1]) + (oa[2] + oa[3]);
        xo[xw * 64 + 2 * lane + 1] = (ob[0] + ob[1]) + (ob[2] + ob[3]);
        xsync();
        if (xw == 0) {
          const float p_x = ex2f(fmaf(xs[8], a.scale_log2e, -mb));
          const float sum = ((xs[4] + xs[5]) + (xs[6] + xs[7])) + p_x;
          const float2 vxf = __half22float2(*reinterpret_cast<const __half2*>(sX + vb * 384 + 128 + lane * 4));
          float o0 = (xo[2 * lane] + xo[64 + 2 * lane]) + (xo[128 + 2 * lane] + xo[192 + 2 * lane]);
          float o1 = (xo[2 * lane + 1] + xo[64 + 2 * lane + 1]) + (xo[128 + 2 * lane + 1] + xo[192 + 2 * lane + 1]);
          o0 = fmaf(p_x, vxf.x, o0), o1 = fmaf(p_x, vxf.y, o1);
          const float inv = 1.0f / sum;
          *reinterpret_cast<uint32_t*>(a.out + (row0 + 256) * hidden + h * 64 + 2 * lane) = pack2(o0 * inv, o1 * inv);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_free[vb]);
      if (xtr) a.trace[(2 * 16 + it) * 16 + 4] = clock64();
    }
  } else {  // ===== warps 4..19 softmax: thread = (query row, half of the key columns)
    const int sw = warp - 4, g = sw >> 3, hf = (sw >> 2) & 1, q = warp & 3;
    const int r = q * 32 + lane, row = g * 128 + r;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 256);
    const uint8_t* q_row = sQ + g * kTileBytes + r * 128;
    uint8_t* prow = sP + (g * 2 + hf) * kTileBytes + r * 128;
    float* my_max = ex_max + (g * 2 + hf) * 128 + r;
    float* my_sum = ex_sum + (g * 2 + hf) * 128 + r;
    const float* other_max = ex_max + (g * 2 + (hf ^ 1)) * 128 + r;
    const float* other_sum = ex_sum + (g * 2 + (hf ^ 1)) * 128 + r;
    int it = 0;
    const bool tracing = a.trace && blockIdx.x == 0 && hf == 0 && q == 0 && lane == 0;  // warps 4 (g0) and 12 (g1)
#define CB_TRACE2(slot, ev)                                                             \
  do {                                                                                  \
    if (tracing && it < 16) a.trace[((slot) * 16 + it) * 16 + (ev)] = clock64();        \
  } while (0)
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, ++it) {
      const int img = u / a.heads, h = u - img * a.heads, xb = it & 1;
      const size_t row0 = (size_t)img * T;
      float s_x = -INFINITY;
      CB_TRACE2(g, 0);
      if (hf == 0) {
        mbar_wait_parked(&q_full[g], it & 1);
        if (has_extra) {  // score against the extra key (token 256): q_row . k_256
          mbar_wait_parked(k_full, it & 1);
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 qa = *reinterpret_cast<const uint4*>(q_row + ((j ^ (r & 7)) << 4));
            const uint4 kxj = *reinterpret_cast<const uint4*>(sX + xb * 384 + j * 16);
            const __half2* q2 = reinterpret_cast<const __half2*>(&qa);
            const __half2* k2 = reinterpret_cast<const __half2*>(&kxj);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 qf = __half22float2(q2[e]), kf = __half22float2(k2[e]);
              acc0 = fmaf(qf.x, kf.x, acc0), acc1 = fmaf(qf.y, kf.y, acc1);
            }
          }
          s_x = acc0 + acc1;
          ex_sx[g * 128 + r] = s_x;
        }
      }
      CB_TRACE2(g, 1);
      mbar_wait_parked(&s_ready[g], it & 1);
      tc_fence_after();
      CB_TRACE2(g, 2);
      if (hf == 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_free[g]), mbar_arrive(k_free);
      }

      // pass 1: maximum over this thread's 128 columns
      uint32_t v[32];
      float mx = s_x;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int k0 = hf * 128 + cc * 32;
        tmem_ld_32x32b_x32(t_row + (uint32_t)k0, v);
        tmem_ld_wait_pin(v);
        if (FULL || k0 + 32 <= t_mma) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + i < t_mma) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      *my_max = mx;
      CB_TRACE2(g, 3);
      group_sync(g);
      CB_TRACE2(g, 4);
      mx = fmaxf(mx, *other_max);
      if (hf == 1 && has_extra) s_x = ex_sx[g * 128 + r];
      const float mb = mx * a.scale_log2e;
      const float p_x = has_extra ? ex2f(fmaf(s_x, a.scale_log2e, -mb)) : 0.f;
      float sum = hf == 0 ? p_x : 0.f, sum1 = 0.f;

      // pass 2: this half-row's two 64-key chunks -> fp16 P in its own swizzled buffer
#pragma unroll 1
      for (int i2 = 0; i2 < 2; ++i2) {
        const int c = hf * 2 + i2, use = it * 2 + i2;
        mbar_wait_parked(&p_free[g * 2 + hf], (use & 1) ^ 1);
        CB_TRACE2(g, 5 + 2 * i2);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int k0 = c * 64 + half * 32;
          tmem_ld_32x32b_x32(t_row + (uint32_t)k0, v);
          tmem_ld_wait_pin(v);
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2f(fmaf(__uint_as_float(v[i]), a.scale_log2e, -mb)), p1 = ex2f(fmaf(__uint_as_float(v[i + 1]), a.scale_log2e, -mb));
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
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g * 2 + hf]);
        CB_TRACE2(g, 6 + 2 * i2);
      }
      sum += sum1;
      *my_sum = sum;
      group_sync(g);
      CB_TRACE2(g, 9);
      sum += *other_sum;

      // epilogue: this thread's 32 output dims
      mbar_wait_parked(&o_ready[g], it & 1);
      if (has_extra) mbar_wait_parked(&v_full[xb], (it >> 1) & 1);
      tc_fence_after();
      CB_TRACE2(g, 10);
      const float inv = 1.0f / sum;
      __half* orow = a.out + (row0 + row) * hidden + h * 64 + hf * 32;
      tmem_ld_32x32b_x32(t_row + (uint32_t)(hf * 32), v);
      tmem_ld_wait_pin(v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * j + e]);
        if (has_extra) {
          const uint4 vxj = *reinterpret_cast<const uint4*>(sX + xb * 384 + 128 + (hf * 4 + j) * 16);
          const __half2* v2 = reinterpret_cast<const __half2*>(&vxj);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 vf = __half22float2(v2[e]);
            o[2 * e] = fmaf(p_x, vf.x, o[2 * e]), o[2 * e + 1] = fmaf(p_x, vf.y, o[2 * e + 1]);
          }
        }
        if (FULL || row < t_mma)
          *reinterpret_cast<uint4*>(orow + j * 8) =
              make_uint4(pack2(o[0] * inv, o[1] * inv), pack2(o[2] * inv, o[3] * inv), pack2(o[4] * inv, o[5] * inv), pack2(o[6] * inv, o[7] * inv));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]), mbar_arrive(&v_free[xb]);
      CB_TRACE2(g, 11);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc2

int attention_tc2(cb_ctx* ctx, const void* qkv, void* out, int n, int tokens, int heads, int head_dim, cudaStream_t stream, bool* launched) {
  *launched = false;
  const char* sel = std::getenv("CB_ATTN_KERNEL");
  if (sel && (std::strcmp(sel, "mma") == 0 || std::strcmp(sel, "tc1") == 0)) return CB_OK;  // A/B switches: tc1 = one thread per row, mma = mma.sync
  if (head_dim != 64 || tokens < 129 || tokens > 257) return CB_OK;
  const int hidden = heads * 64;
  CUtensorMap map, map_row;
  const uint64_t dims[2] = {(uint64_t)3 * hidden, (uint64_t)n * tokens}, strides[1] = {(uint64_t)3 * hidden * 2};
  const uint32_t box[2] = {64, 128}, box_row[2] = {64, 1};
  int rc = make_tensor_map(ctx, &map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map(ctx, &map_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qkv, dims, strides, box_row, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  static bool attr_done[64] = {};  // the attribute is per device: one process may drive several
  bool& attr_set = attr_done[ctx->device & 63];
  if (!attr_set) {
    CB_CUDA(ctx, cudaFuncSetAttribute(tc2::attention_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmem));
    CB_CUDA(ctx, cudaFuncSetAttribute(tc2::attention_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmem));
    attr_set = true;
  }
  tc2::Args a{(const __half*)qkv, (__half*)out, tokens, heads, n * heads, 1.4426950408889634f / sqrtf(64.f), nullptr};
  const char* trc = std::getenv("CB_ATTN_DEBUG_TRACE");
  const bool tracing = trc && trc[0] == '1';
  if (tracing) {
    CB_CUDA(ctx, cudaMalloc(&a.trace, 3 * 16 * 16 * sizeof(long long)));
    CB_CUDA(ctx, cudaMemset(a.trace, 0, 3 * 16 * 16 * sizeof(long long)));
  }
  const int grid = std::min(n * heads, ctx->sm_count);
  mark_launch(ctx, CB_PROF_ATTENTION, stream);
  if (tokens >= 256)
    tc2::attention_tc2_kernel<true><<<grid, tc2::kThreads, tc2::kSmem, stream>>>(map, map_row, a);
  else
    tc2::attention_tc2_kernel<false><<<grid, tc2::kThreads, tc2::kSmem, stream>>>(map, map_row, a);
  CB_CUDA(ctx, cudaGetLastError());
  if (tracing) {
    long long hbuf[3 * 16 * 16];
    CB_CUDA(ctx, cudaStreamSynchronize(stream));
    CB_CUDA(ctx, cudaMemcpy(hbuf, a.trace, sizeof(hbuf), cudaMemcpyDeviceToHost));
    cudaFree(a.trace);
    const long long t0 = hbuf[0];
    for (int s = 0; s < 3; ++s)
      for (int it = 2; it < 7; ++it) {
        printf("trace2 %s it%d:", s == 0 ? "g0" : (s == 1 ? "g1" : "xr"), it);
        for (int e = 0; e < 12; ++e) printf(" %lld", hbuf[(s * 16 + it) * 16 + e] ? hbuf[(s * 16 + it) * 16 + e] - t0 : -1);
        printf("\n");
      }
  }
  *launched = true;
  return CB_OK;
}

}  // namespace cb
