// Backward of the fused HGT edge kernel (see edge.cu for the forward and the lane mapping).
//
// Forward per destination i and head h:  s_e = <q_i, k_e>,  p_e = exp(s_e - m_i) / (l_i + 1e-16),
//                                         agg_i = sum_e p_e v_e            (k_e / v_e include the RTE rows)
// Given dagg_i:   dv_e = p_e * dagg_i          dp_e = <dagg_i, v_e>          D_i = <dagg_i, agg_i> (= sum_e p_e dp_e)
//                 ds_e = p_e * (dp_e - D_i)    dq_i = sum_e ds_e k_e         dk_e = ds_e * q_i
// One pass over the destination-sorted CSR; p is recomputed from the saved per-(destination, head) (m, l).
// dq is owned by the destination's warp (atomics only for split hub pieces); dk / dv are scattered into the
// [K'|V'] gradient table (and the RTE gradient table) with vector fp32 reductions (red.global.add.v4.f32):
// many edges share a <source, relation> row.  The trailing all-zero row collects the gradient of edges that
// matched no triple and is discarded by the caller.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;

struct BwdParams {
  const float* q;
  const float* kv;
  const float* kvr;
  const float* agg;
  const float* dagg;
  const float* stats;        // [N, 2H] (m, l)
  const int32_t* row_ptr;
  const int32_t* kv_row;
  const int32_t* rte_row;
  const int32_t* tiles;
  int32_t n_tiles;
  const int32_t* d_counts;   // optional device {n_tiles, ...} (sync-free plans): n_tiles is then an upper bound
  int32_t d, H, DK, LPH, lph_shift;
  float* dq;                 // [N, d]   zero-initialised by hgt_edge_backward
  float* dkv;                // [rows+1, 2d] zero-initialised
  float* dkvr;               // [P*240+1, 2d] zero-initialised or nullptr
  int32_t* tile_counter;
};

template <int VEC>
__device__ __forceinline__ void ld_vec(float (&dst)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    dst[0] = t.x; dst[1] = t.y;
  } else {
    dst[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void red_add_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
  } else if constexpr (VEC == 2) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
  } else {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
  }
}
__device__ __forceinline__ float head_sum(float v, int lph) {
  for (int o = lph >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VEC, int NCH>
__global__ void __launch_bounds__(kWarps * 32)
k_edge_bwd(BwdParams p) {
  const int lane = threadIdx.x & 31;
  const int lph = p.LPH;
  const int h = lane >> p.lph_shift;
  const int sub = lane & (lph - 1);
  const bool head_ok = h < p.H;
  int offs[NCH];
#pragma unroll
  for (int t = 0; t < NCH; ++t) {
    int o = (sub + t * lph) * VEC;
    offs[t] = (head_ok && o < p.DK) ? h * p.DK + o : -1;
  }
  const int64_t row_stride = 2 * (int64_t)p.d;
  const bool rte = p.kvr != nullptr;

  const int n_tiles = p.d_counts ? p.d_counts[0] : p.n_tiles;
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(p.tile_counter, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    const int4 tl = reinterpret_cast<const int4*>(p.tiles)[tile];
    const bool split = tl.y < 0;
    const int d_begin = tl.x, d_end = split ? tl.x + 1 : tl.y;
    int seg_begin = tl.z;
    for (int dst = d_begin; dst < d_end; ++dst) {
      const int seg_end = split ? tl.w : p.row_ptr[dst + 1];
      if (seg_end > seg_begin) {
        float q[NCH][VEC], da[NCH][VEC], dq[NCH][VEC];
        float dpart = 0.f;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (offs[t] >= 0) {
            float ag[VEC];
            ld_vec<VEC>(q[t], p.q + (int64_t)dst * p.d + offs[t]);
            ld_vec<VEC>(da[t], p.dagg + (int64_t)dst * p.d + offs[t]);
            ld_vec<VEC>(ag, p.agg + (int64_t)dst * p.d + offs[t]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) dpart = fmaf(da[t][v], ag[v], dpart);
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { q[t][v] = 0.f; da[t][v] = 0.f; }
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) dq[t][v] = 0.f;
        }
        const float D = head_sum(dpart, lph);
        float m = 0.f, inv_l = 0.f;
        if (head_ok) {
          m = p.stats[(int64_t)dst * 2 * p.H + h];
          inv_l = 1.0f / (p.stats[(int64_t)dst * 2 * p.H + p.H + h] + 1e-16f);
        }
        for (int c = seg_begin; c < seg_end; ++c) {
          const int64_t row = p.kv_row[c];
          const int64_t rrow = rte ? p.rte_row[c] : 0;
          const float* kvp = p.kv + row * row_stride;
          float kk[NCH][VEC], vv[NCH][VEC];
          float spart = 0.f, dppart = 0.f;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (offs[t] >= 0) {
              ld_vec<VEC>(kk[t], kvp + offs[t]);
              ld_vec<VEC>(vv[t], kvp + p.d + offs[t]);
              if (rte) {
                float a[VEC], b[VEC];
                ld_vec<VEC>(a, p.kvr + rrow * row_stride + offs[t]);
                ld_vec<VEC>(b, p.kvr + rrow * row_stride + p.d + offs[t]);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { kk[t][v] += a[v]; vv[t][v] += b[v]; }
              }
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                spart = fmaf(q[t][v], kk[t][v], spart);
                dppart = fmaf(da[t][v], vv[t][v], dppart);
              }
            }
          }
          const float s = head_sum(spart, lph);
          const float dp = head_sum(dppart, lph);
          const float pe = __expf(s - m) * inv_l;
          const float ds = pe * (dp - D);
          float* gk = p.dkv + row * row_stride;
#pragma unroll
          for (int t = 0; t < NCH; ++t) {
            if (offs[t] >= 0) {
              float gkv[VEC], gvv[VEC];
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                dq[t][v] = fmaf(ds, kk[t][v], dq[t][v]);
                gkv[v] = ds * q[t][v];
                gvv[v] = pe * da[t][v];
              }
              red_add_vec<VEC>(gk + offs[t], gkv);
              red_add_vec<VEC>(gk + p.d + offs[t], gvv);
              if (rte) {
                red_add_vec<VEC>(p.dkvr + rrow * row_stride + offs[t], gkv);
                red_add_vec<VEC>(p.dkvr + rrow * row_stride + p.d + offs[t], gvv);
              }
            }
          }
        }
        float* gq = p.dq + (int64_t)dst * p.d;
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (offs[t] >= 0) {
            if (split) red_add_vec<VEC>(gq + offs[t], dq[t]);
            else {
#pragma unroll
              for (int v = 0; v < VEC; ++v) gq[offs[t] + v] = dq[t][v];
            }
          }
        }
      }
      seg_begin = seg_end;
    }
  }
}

template <int VEC>
int dispatch(const BwdParams& p, int nch, int grid, cudaStream_t st) {
  switch (nch) {
    case 1: k_edge_bwd<VEC, 1><<<grid, kWarps * 32, 0, st>>>(p); break;
    case 2: k_edge_bwd<VEC, 2><<<grid, kWarps * 32, 0, st>>>(p); break;
    case 4: k_edge_bwd<VEC, 4><<<grid, kWarps * 32, 0, st>>>(p); break;
    case 8: k_edge_bwd<VEC, 8><<<grid, kWarps * 32, 0, st>>>(p); break;
    default: hgt_set_error("hgt_edge_backward: unsupported chunk count %d", nch); return 1;
  }
  HGT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int hgt_edge_backward(const float* q, const float* kv, const float* kvr, const float* agg,
                                 const float* dagg, const float* stats, const int32_t* row_ptr,
                                 const int32_t* kv_row, const int32_t* rte_row, const int32_t* tiles, int32_t n_tiles,
                                 int64_t n_nodes, int32_t d, int32_t n_heads, int64_t kv_rows_total,
                                 int64_t kvr_rows_total, float* dq, float* dkv, float* dkvr,
                                 void* workspace, size_t workspace_bytes, const int32_t* d_tile_counts, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  HGT_REQUIRE(n_heads >= 1 && n_heads <= 32 && d % n_heads == 0, "hgt_edge_backward: bad d=%d / n_heads=%d", d, n_heads);
  HGT_REQUIRE((kvr != nullptr) == (rte_row != nullptr) && (kvr != nullptr) == (dkvr != nullptr),
              "hgt_edge_backward: kvr, rte_row and dkvr must go together");
  HGT_REQUIRE(workspace && workspace_bytes >= 256, "hgt_edge_backward: workspace too small");
  // the kernel accumulates (dk / dv of a <source, relation> row come from many edges): this call owns the initialisation
  if (n_nodes > 0) HGT_CHECK_CUDA(cudaMemsetAsync(dq, 0, (size_t)n_nodes * d * sizeof(float), st));
  if (kv_rows_total > 0) HGT_CHECK_CUDA(cudaMemsetAsync(dkv, 0, (size_t)kv_rows_total * 2 * d * sizeof(float), st));
  if (dkvr && kvr_rows_total > 0)
    HGT_CHECK_CUDA(cudaMemsetAsync(dkvr, 0, (size_t)kvr_rows_total * 2 * d * sizeof(float), st));
  if (n_nodes == 0 || n_tiles == 0) return 0;
  BwdParams p;
  p.q = q; p.kv = kv; p.kvr = kvr; p.agg = agg; p.dagg = dagg; p.stats = stats; p.row_ptr = row_ptr;
  p.kv_row = kv_row; p.rte_row = rte_row; p.tiles = tiles; p.n_tiles = n_tiles; p.d_counts = d_tile_counts; p.d = d; p.H = n_heads;
  p.DK = d / n_heads;
  int hp = 1;
  while (hp < n_heads) hp <<= 1;
  p.LPH = 32 / hp;
  int shift = 0;
  while ((1 << shift) < p.LPH) ++shift;
  p.lph_shift = shift;
  p.dq = dq; p.dkv = dkv; p.dkvr = dkvr;
  p.tile_counter = reinterpret_cast<int32_t*>(workspace);
  int vec = 1;
  for (int v : {4, 2})
    if (p.DK % v == 0 && p.DK / v >= p.LPH) { vec = v; break; }
  int chunks = (p.DK + vec * p.LPH - 1) / (vec * p.LPH), nch = 1;
  while (nch < chunks) nch <<= 1;
  HGT_REQUIRE(nch <= 8, "hgt_edge_backward: head width d_k=%d needs %d chunks per lane (max 8)", p.DK, chunks);
  HGT_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int32_t), st));
  int grid = hgt_sm_count() * 4;
  int max_ctas = (n_tiles + kWarps - 1) / kWarps;
  if (grid > max_ctas) grid = max_ctas;
  if (vec == 4) return dispatch<4>(p, nch, grid, st);
  if (vec == 2) return dispatch<2>(p, nch, grid, st);
  return dispatch<1>(p, nch, grid, st);
}
