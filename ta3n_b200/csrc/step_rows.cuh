// step_rows.cuh -- the row ("SIMT") tasks of the fused training step: everything of the path that is not a
// dense contraction, written as device functions over a group of 256 threads so that they run either as
// stand-alone kernels (phased executor) or as tasks of the persistent step kernel (step_kernel.cuh).
//
//   tail_task   : per video, everything between the relation-discriminator hidden layer and the data-gradient
//                 GEMMs -- models.py:479 (relation heads), :351-357 + :379-388 + :651-652 (entropy attention,
//                 attentive pooling), :679-687 (dropout + classifier), :464-470 (video discriminator, both layers),
//                 :456-462 (frame head), main.py:446, 508-538, 559-562 + loss.py:15-25 (all loss heads and their
//                 gradients), then the backward of the same ops down to the operands of the dgrad GEMMs.
//                 Replaces ten launches of the per-op sequence (relattn_fwd, head_fwd x3, two tensor-core launches
//                 of the video discriminator, loss_heads, loss_reduce, head_bwd_data, video_head_bwd,
//                 relattn_bwd_pre).
//   colsum_task : one (job, column block, row split) of the deterministic weighted column sums (bias gradients and
//                 the skinny head weight gradients), and the fixed-order reduction of its row splits.
#pragma once

#include "rowops.cuh"

namespace ta3n {

constexpr int kRowThreads = 256;          // threads of a row task (8 warps)
constexpr int kTailVideos = 4;            // videos per tail task
constexpr int kTailMaxT = 32, kTailMaxC = 128;

__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct TailArgs {
  int M, Bs, T, R, H, F, C, n_rel;
  int use_attn, loss_flags;
  float gamma;
  float dom_w0, dom_w1;              // weight of criterion_domain (main.py:165-167)
  const float* class_weight;         // [C] weight of criterion (main.py:160-163) or nullptr
  const float* beta_dev;             // [3] relation, video, frame GRL coefficients (device: main.py:350-352)
  const long long* labels;
  const int* valid_rows;
  RelMap map;
  // forward inputs
  const float* hid_f;                // [M*T, F]
  const float* act;                  // [n_rel, M, H]
  const float* hid_r;                // [R, M, H]
  // weights
  const float* W2f;
  const float* b2f;
  PtrTable W2r, b2r;
  const float* Wc;
  const float* bc;
  const float* W1v;
  const float* b1v;
  const float* W2v;
  const float* b2v;
  DropArgs drop_v;
  // forward outputs
  float* pred_frame;                 // [M*T, 2]
  float* feat_rel;                   // [M, R, H]
  float* pred_rel;                   // [M, R, 2]
  float* attn;                       // [M, R]
  float* feat_video;                 // [M, H]
  float* dropped;                    // [M, H]
  float* pred_video;                 // [M, C]
  float* hid_v;                      // [M, H]
  float* pred_dom;                   // [M, 2]
  float* row_loss;                   // [M]
  // backward outputs (operands of the dgrad / wgrad GEMMs and of the column sums)
  float* g_video;                    // [M, C]
  float* g_dom;                      // [M, 2]
  float* g_frame;                    // [M*T, 2]
  float* Pt;                         // [M, R, 2]
  float* dHv;                        // [M, H]
  float* G;                          // [M, H]     d loss / d feat_video
  float* dHid;                       // [R, M, H]
  float* dHf;                        // [M*T, F]
  unsigned long long* dbg;           // optional [tasks][16] phase timestamps (development)
};

// dot of a row held in shared memory (len floats) with a global row, distributed over a warp
__device__ __forceinline__ float warp_dot_sg(const float* __restrict__ s, const float* __restrict__ g, int len,
                                             int lane) {
  float acc = 0.f;
  for (int k = lane * 4; k + 3 < len; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(s + k);
    const float4 b = __ldg(reinterpret_cast<const float4*>(g + k));
    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
  }
  return warp_sum(acc);
}

// One tail task: videos [v0, v0 + nv), nv <= kTailVideos.  `sm` is >= tail_smem_floats() floats of shared memory,
// 16-byte aligned; tid in [0, 256).  Requires H % 128 == 0 (H = 256 on the path), F % 4 == 0.
// shared memory (floats) of a task of `vp` videos:
// feat_rel [vp][R][H] | 5 vectors [vp][H] (dropped, hv, dHv, G, fv) | logits and their gradients
__host__ __device__ constexpr int tail_smem_floats(int vp, int R, int H, int T, int C) {
  return vp * R * H + 5 * vp * H + vp * (2 * (2 * T + 2 * R + C + 2) + R) + 16;
}

// vp = videos the shared-memory layout is sized for (nv <= vp <= kTailVideos)
// NOTE: force-inlined, and every scalar / pointer of the argument block is copied into a local first.  As an
// out-of-line function taking `const TailArgs&`, each field access was a GENERIC load (the block lives in parameter
// or shared memory) that the compiler had to repeat after every store (possible aliasing): the task took 95 us.
struct TailLocal {      // the pointer / scalar fields of TailArgs, as restrict-qualified locals
  const float* __restrict__ hid_f;
  const float* __restrict__ act;
  const float* __restrict__ hid_r;
  const float* __restrict__ W2f;
  const float* __restrict__ b2f;
  const float* __restrict__ Wc;
  const float* __restrict__ bc;
  const float* __restrict__ W1v;
  const float* __restrict__ b1v;
  const float* __restrict__ W2v;
  const float* __restrict__ b2v;
  const float* __restrict__ class_weight;
  const long long* __restrict__ labels;
  float* __restrict__ pred_frame;
  float* __restrict__ feat_rel;
  float* __restrict__ pred_rel;
  float* __restrict__ attn;
  float* __restrict__ feat_video;
  float* __restrict__ dropped;
  float* __restrict__ pred_video;
  float* __restrict__ hid_v;
  float* __restrict__ pred_dom;
  float* __restrict__ row_loss;
  float* __restrict__ g_video;
  float* __restrict__ g_dom;
  float* __restrict__ g_frame;
  float* __restrict__ Pt;
  float* __restrict__ dHv;
  float* __restrict__ G;
  float* __restrict__ dHid;
  float* __restrict__ dHf;
  DropArgs drop_v;
  int use_attn, loss_flags, Bs;
  float gamma, dom_w0, dom_w1;
};

__device__ __forceinline__ void tail_task(const TailArgs& args, const int v0, const int nv, const int vp,
                                          float* __restrict__ sm, const int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int M = args.M, T = args.T, R = args.R, H = args.H, F = args.F, C = args.C;
  TailLocal a;
  a.hid_f = args.hid_f; a.act = args.act; a.hid_r = args.hid_r; a.W2f = args.W2f; a.b2f = args.b2f; a.Wc = args.Wc;
  a.bc = args.bc; a.W1v = args.W1v; a.b1v = args.b1v; a.W2v = args.W2v; a.b2v = args.b2v;
  a.class_weight = args.class_weight; a.labels = args.labels; a.pred_frame = args.pred_frame;
  a.feat_rel = args.feat_rel; a.pred_rel = args.pred_rel; a.attn = args.attn; a.feat_video = args.feat_video;
  a.dropped = args.dropped; a.pred_video = args.pred_video; a.hid_v = args.hid_v; a.pred_dom = args.pred_dom;
  a.row_loss = args.row_loss; a.g_video = args.g_video; a.g_dom = args.g_dom; a.g_frame = args.g_frame; a.Pt = args.Pt;
  a.dHv = args.dHv; a.G = args.G; a.dHid = args.dHid; a.dHf = args.dHf; a.drop_v = args.drop_v;
  a.use_attn = args.use_attn; a.loss_flags = args.loss_flags; a.Bs = args.Bs; a.gamma = args.gamma;
  a.dom_w0 = args.dom_w0; a.dom_w1 = args.dom_w1;
  const int vs_in = args.valid_rows ? args.valid_rows[0] : args.Bs;
  const int vt_in = args.valid_rows ? args.valid_rows[1] : args.M - args.Bs;
  unsigned long long* const dbg = args.dbg;
  constexpr int V = kTailVideos;                  // register arrays; smem strides use vp
  float* s_fr = sm;                               // [vp][R][H]  feat_rel
  float* s_drop = s_fr + vp * R * H;              // [vp][H]     dropped features
  float* s_hv = s_drop + vp * H;                  // [vp][H]     video-disc hidden
  float* s_dhv = s_hv + vp * H;                   // [vp][H]
  float* s_G = s_dhv + vp * H;                    // [vp][H]
  float* s_fv = s_G + vp * H;                     // [vp][H]     feat_video (pre-dropout)
  float* s_pf = s_fv + vp * H;                    // [vp][T][2]  frame logits
  float* s_pr = s_pf + vp * 2 * T;                // [vp][R][2]  relation logits
  float* s_pv = s_pr + vp * 2 * R;                // [vp][C]     class logits
  float* s_pd = s_pv + vp * C;                    // [vp][2]     video-domain logits
  float* s_gf = s_pd + vp * 2;                    // [vp][T][2]  gradients of the above
  float* s_gr = s_gf + vp * 2 * T;
  float* s_gv = s_gr + vp * 2 * R;
  float* s_gd = s_gv + vp * C;
  float* s_w = s_gd + vp * 2;                     // [vp][R]     attention weight + 1
  const float beta1 = args.beta_dev ? __ldg(args.beta_dev + 1) : 0.f;
  int dbg_i = 0;
#define TAIL_MARK()                                                                          \
  do {                                                                                       \
    if (dbg && tid == 0) {                                                                   \
      unsigned long long t_;                                                                 \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                  \
      dbg[(size_t)(v0 / kTailVideos) * 16 + dbg_i] = t_;                                     \
    }                                                                                        \
    ++dbg_i;                                                                                 \
  } while (0)
  TAIL_MARK();

  // ---- phase 1: frame logits, relation sums + relation logits ---------------------------------------
  for (int it = warp; it < nv * T; it += 8) {               // warp per frame row: 2 dots of length F
    const int v = it / T, t = it - v * T;
    const float* hr = a.hid_f + ((size_t)(v0 + v) * T + t) * F;
    float s0 = 0.f, s1 = 0.f;
    for (int k = lane * 4; k < F; k += 128) {
      const float4 h = __ldcg(reinterpret_cast<const float4*>(hr + k));
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.W2f + k));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.W2f + F + k));
      s0 = fmaf(h.x, w0.x, fmaf(h.y, w0.y, fmaf(h.z, w0.z, fmaf(h.w, w0.w, s0))));
      s1 = fmaf(h.x, w1.x, fmaf(h.y, w1.y, fmaf(h.z, w1.z, fmaf(h.w, w1.w, s1))));
    }
    s0 = warp_sum(s0) + __ldg(a.b2f);
    s1 = warp_sum(s1) + __ldg(a.b2f + 1);
    if (lane == 0) {
      s_pf[(v * T + t) * 2] = s0;
      s_pf[(v * T + t) * 2 + 1] = s1;
      a.pred_frame[((size_t)(v0 + v) * T + t) * 2] = s0;
      a.pred_frame[((size_t)(v0 + v) * T + t) * 2 + 1] = s1;
    }
  }
  {  // feat_rel[v,i,:] = sum_r act[q(i,r)][v,:]                                       TRNmodule.py:79
    const int H4 = H >> 2;
    const size_t plane = (size_t)M * H;
    for (int e = tid; e < nv * R * H4; e += kRowThreads) {
      const int h4 = e % H4, vi = e / H4, i = vi % R, v = vi / R;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = args.map.rel_begin[i]; q < args.map.rel_begin[i + 1]; ++q) {
        const float4 x = __ldcg(reinterpret_cast<const float4*>(a.act + q * plane + (size_t)(v0 + v) * H) + h4);
        s.x += x.x;
        s.y += x.y;
        s.z += x.z;
        s.w += x.w;
      }
      reinterpret_cast<float4*>(s_fr + (v * R + i) * H)[h4] = s;
      reinterpret_cast<float4*>(a.feat_rel + ((size_t)(v0 + v) * R + i) * H)[h4] = s;
    }
  }
  for (int it = warp; it < nv * R; it += 8) {               // relation logits: hid_r[i][v,:] . W2r_i  (models.py:479)
    const int v = it / R, i = it - v * R;
    const float* hr = a.hid_r + ((size_t)i * M + (v0 + v)) * H;
    const float* w0 = args.W2r.p[i];
    float s0 = 0.f, s1 = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 h = __ldcg(reinterpret_cast<const float4*>(hr + k));
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(w0 + k));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(w0 + H + k));
      s0 = fmaf(h.x, x0.x, fmaf(h.y, x0.y, fmaf(h.z, x0.z, fmaf(h.w, x0.w, s0))));
      s1 = fmaf(h.x, x1.x, fmaf(h.y, x1.y, fmaf(h.z, x1.z, fmaf(h.w, x1.w, s1))));
    }
    s0 = warp_sum(s0) + __ldg(args.b2r.p[i]);
    s1 = warp_sum(s1) + __ldg(args.b2r.p[i] + 1);
    if (lane == 0) {
      const size_t o = (size_t)(v0 + v) * R + i;
      s_pr[(v * R + i) * 2] = s0;
      s_pr[(v * R + i) * 2 + 1] = s1;
      a.pred_rel[o * 2] = s0;
      a.pred_rel[o * 2 + 1] = s1;
      const float w = a.use_attn ? attn_from_logits(s0, s1).w : 0.f;          // models.py:351-357
      s_w[v * R + i] = w + 1.0f;
      if (a.use_attn) a.attn[o] = w;
    }
  }
  row_sync();
  TAIL_MARK();
  // ---- phase 2: attentive pooling + dropout                         models.py:379-388, 651-652, 679-680 ----
  for (int e = tid; e < nv * H; e += kRowThreads) {
    const int v = e / H, h = e - v * H;
    float y = 0.f;
    for (int i = 0; i < R; ++i) y = fmaf(s_w[v * R + i], s_fr[(v * R + i) * H + h], y);
    const size_t ge = (size_t)(v0 + v) * H + h;
    a.feat_video[ge] = y;
    s_fv[e] = y;
    const float d = y * drop_factor(a.drop_v, ge);
    s_drop[e] = d;
    a.dropped[ge] = d;
  }
  if (!a.use_attn)                                          // models.py:647 placeholder output
    for (int e = tid; e < nv * R; e += kRowThreads) a.attn[(size_t)v0 * R + e] = s_fr[e * H];
  row_sync();
  TAIL_MARK();
  // ---- phase 3: classifier logits, video-discriminator hidden layer        models.py:681-687, 464-468 ----
  // Both are latency-bound walks over weights in L2: every warp issues the loads of a whole batch of rows before it
  // touches the first one (8 rows = 16 independent 16 B loads per lane in flight).
  {
    const int n_items = nv * C;
    for (int base = warp; base < n_items; base += 8 * 8) {
      float4 w[8][2];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int it = base + 8 * u;
        const int c = it < n_items ? it % C : 0;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = lane * 4 + 128 * kk;
          w[u][kk] = (it < n_items && k < H) ? __ldg(reinterpret_cast<const float4*>(a.Wc + (size_t)c * H + k))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int it = base + 8 * u;
        if (it >= n_items) break;
        const int v = it / C, c = it - v * C;
        float acc = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = lane * 4 + 128 * kk;
          if (k < H) {
            const float4 d = *reinterpret_cast<const float4*>(s_drop + v * H + k);
            acc = fmaf(d.x, w[u][kk].x, fmaf(d.y, w[u][kk].y, fmaf(d.z, w[u][kk].z, fmaf(d.w, w[u][kk].w, acc))));
          }
        }
        for (int k = lane * 4 + 256; k < H; k += 128) {          // H > 256: the rest of the row, plainly
          const float4 d = *reinterpret_cast<const float4*>(s_drop + v * H + k);
          const float4 x = __ldg(reinterpret_cast<const float4*>(a.Wc + (size_t)c * H + k));
          acc = fmaf(d.x, x.x, fmaf(d.y, x.y, fmaf(d.z, x.z, fmaf(d.w, x.w, acc))));
        }
        const float sres = warp_sum(acc) + __ldg(a.bc + c);
        if (lane == 0) {
          s_pv[v * C + c] = sres;
          a.pred_video[(size_t)(v0 + v) * C + c] = sres;
        }
      }
    }
  }
  for (int h0 = warp * (H / 8); h0 < (warp + 1) * (H / 8); h0 += 8) {     // warp w owns hidden units [w H/8, (w+1) H/8)
    float4 w[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = lane * 4 + 128 * kk;
        w[u][kk] = k < H ? __ldg(reinterpret_cast<const float4*>(a.W1v + (size_t)(h0 + u) * H + k))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int h = h0 + u;
      float acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int k = lane * 4 + 128 * kk;
        if (k < H) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nv) {
              const float4 d = *reinterpret_cast<const float4*>(s_drop + v * H + k);
              acc[v] = fmaf(d.x, w[u][kk].x, fmaf(d.y, w[u][kk].y, fmaf(d.z, w[u][kk].z, fmaf(d.w, w[u][kk].w, acc[v]))));
            }
        }
      }
      for (int k = lane * 4 + 256; k < H; k += 128) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a.W1v + (size_t)h * H + k));
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v < nv) {
            const float4 d = *reinterpret_cast<const float4*>(s_drop + v * H + k);
            acc[v] = fmaf(d.x, x.x, fmaf(d.y, x.y, fmaf(d.z, x.z, fmaf(d.w, x.w, acc[v]))));
          }
      }
      const float b = __ldg(a.b1v + h);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float sres = fmaxf(warp_sum(acc[v]) + b, 0.f);
        if (lane == 0 && v < nv) {
          s_hv[v * H + h] = sres;
          a.hid_v[(size_t)(v0 + v) * H + h] = sres;
        }
      }
    }
  }
  row_sync();
  TAIL_MARK();
  // ---- phase 4: video-domain logits                                              models.py:469-470 ----
  for (int it = warp; it < nv * 2; it += 8) {
    const int v = it >> 1, j = it & 1;
    const float s = warp_dot_sg(s_hv + v * H, a.W2v + (size_t)j * H, H, lane) + __ldg(a.b2v + j);
    if (lane == 0) {
      s_pd[v * 2 + j] = s;
      a.pred_dom[(size_t)(v0 + v) * 2 + j] = s;
    }
  }
  row_sync();
  TAIL_MARK();
  // ---- phase 5: loss heads (one warp per video)          main.py:446, 508-538, 559-562; loss.py:15-25 ----
  {
    const int Bs = a.Bs;
    const int vs = min(vs_in, Bs);
    const int vt = min(vt_in, M - Bs);
    // normalisers of the (weighted) means: CrossEntropyLoss(weight=w) divides by the sum of the weights of the rows
    float n_cls = (float)max(vs, 1);
    if (a.class_weight) {                                   // main.py:160-163, 204: sum_m w[y_m] over the real source rows
      float s = 0.f;
      for (int m = lane; m < vs; m += 32) s += __ldg(a.class_weight + (int)a.labels[m]);
      n_cls = fmaxf(warp_sum(s), 1e-30f);
    }
    const float n_dom = fmaxf(a.dom_w0 * (float)vs + a.dom_w1 * (float)vt, 1e-30f);   // per level: times rows per video
    const float n_all = (float)max(vs + vt, 1);
    for (int v = warp; v < nv; v += 8) {
      const int m = v0 + v;
      const int dom = m >= Bs ? 1 : 0;
      float* gv = s_gv + v * C;
      if (dom ? (m - Bs >= vt) : (m >= vs)) {               // padding row of a short last batch (main.py:354-372, 421-422)
        for (int c = lane; c < C; c += 32) gv[c] = 0.f;
        for (int i = lane; i < 2 * R; i += 32) s_gr[v * 2 * R + i] = 0.f;
        for (int t = lane; t < 2 * T; t += 32) s_gf[v * 2 * T + t] = 0.f;
        if (lane < 2) s_gd[v * 2 + lane] = 0.f;
        if (lane == 0) a.row_loss[m] = 0.f;
        continue;
      }
      const float wd = dom ? a.dom_w1 : a.dom_w0;
      const float* pv = s_pv + v * C;
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, pv[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += expf(pv[c] - mx);
      se = warp_sum(se);
      const float lse = logf(se);
      float hc = 0.f;                                       // entropy of the class prediction
      for (int c = lane; c < C; c += 32) {
        const float lq = pv[c] - mx - lse;
        hc -= expf(lq) * lq;
      }
      hc = warp_sum(hc);
      const Attn2 dv = attn_from_logits(s_pd[v * 2], s_pd[v * 2 + 1]);
      const bool att = (a.loss_flags & LOSS_ATT_ENT) != 0;
      const float att_scale = att ? a.gamma / n_all : 0.f;
      const int y = (m < Bs) ? (int)a.labels[m] : -1;
      const float wy = (m < Bs) ? (a.class_weight ? __ldg(a.class_weight + y) : 1.f) : 0.f;
      float loss = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float lq = pv[c] - mx - lse;
        const float q = expf(lq);
        float gq = 0.f;
        if (m < Bs) gq = wy * (q - (c == y ? 1.f : 0.f)) / n_cls;
        gq += att_scale * (1.f + dv.ent) * (-q * (lq + hc));
        gv[c] = gq;
        if (m < Bs && c == y) loss += -wy * lq / n_cls;
      }
      loss = warp_sum(loss);                                // exactly one lane held the CE term
      float l = loss + att_scale * (1.f + dv.ent) * hc;
      float g0 = 0.f, g1 = 0.f;
      if (a.loss_flags & LOSS_ADV_VIDEO) {
        l += -wd * (dom ? dv.lq1 : dv.lq0) / n_dom;
        g0 = wd * (dv.q0 - (dom ? 0.f : 1.f)) / n_dom;
        g1 = wd * (dv.q1 - (dom ? 1.f : 0.f)) / n_dom;
      }
      g0 += att_scale * hc * (-dv.q0 * (dv.lq0 + dv.ent));
      g1 += att_scale * hc * (-dv.q1 * (dv.lq1 + dv.ent));
      if (lane == 0) {
        s_gd[v * 2] = g0;
        s_gd[v * 2 + 1] = g1;
      }
      float extra = 0.f;
      for (int i = lane; i < R; i += 32) {
        float r0 = 0.f, r1 = 0.f;
        if (a.loss_flags & LOSS_ADV_REL) {
          const Attn2 x = attn_from_logits(s_pr[(v * R + i) * 2], s_pr[(v * R + i) * 2 + 1]);
          const float inv = wd / (n_dom * (float)R);
          extra += -(dom ? x.lq1 : x.lq0) * inv;
          r0 = (x.q0 - (dom ? 0.f : 1.f)) * inv;
          r1 = (x.q1 - (dom ? 1.f : 0.f)) * inv;
        }
        s_gr[(v * R + i) * 2] = r0;
        s_gr[(v * R + i) * 2 + 1] = r1;
      }
      for (int t = lane; t < T; t += 32) {
        float f0 = 0.f, f1 = 0.f;
        if (a.loss_flags & LOSS_ADV_FRAME) {
          const Attn2 x = attn_from_logits(s_pf[(v * T + t) * 2], s_pf[(v * T + t) * 2 + 1]);
          const float inv = wd / (n_dom * (float)T);
          extra += -(dom ? x.lq1 : x.lq0) * inv;
          f0 = (x.q0 - (dom ? 0.f : 1.f)) * inv;
          f1 = (x.q1 - (dom ? 1.f : 0.f)) * inv;
        }
        s_gf[(v * T + t) * 2] = f0;
        s_gf[(v * T + t) * 2 + 1] = f1;
      }
      extra = warp_sum(extra);
      if (lane == 0) a.row_loss[m] = l + extra;
    }
  }
  row_sync();
  TAIL_MARK();
  // the head gradients are operands of the column sums (skinny weight gradients and bias gradients)
  for (int e = tid; e < nv * C; e += kRowThreads) a.g_video[(size_t)v0 * C + e] = s_gv[e];
  for (int e = tid; e < nv * 2; e += kRowThreads) a.g_dom[(size_t)v0 * 2 + e] = s_gd[e];
  for (int e = tid; e < nv * 2 * T; e += kRowThreads) a.g_frame[(size_t)v0 * 2 * T + e] = s_gf[e];
  // ---- phase 6: dHv = (g_dom W2v) * 1[hid_v > 0]                                      (head_bwd_data) ----
  for (int e = tid; e < nv * H; e += kRowThreads) {
    const int v = e / H, h = e - v * H;
    const float s = s_hv[e] > 0.f ? fmaf(s_gd[v * 2], __ldg(a.W2v + h), s_gd[v * 2 + 1] * __ldg(a.W2v + H + h)) : 0.f;
    s_dhv[e] = s;
    a.dHv[(size_t)(v0 + v) * H + h] = s;
  }
  row_sync();
  TAIL_MARK();
  // ---- phase 7: G = ((g_video Wc) - beta1 * (dHv W1v)) * keep/(1-p)      (disc dgrad + video_head_bwd) ----
  // d[v, k] = sum_h dHv[v, h] W1v[h, k]: lane = (kq, hg) -- 8 groups of 4 consecutive k per warp, the 4 lanes of a group
  // split the h range, 16 independent 16 B loads in flight per lane, then two shuffles fold the h quarters.
  for (int kbase = 0; kbase < H; kbase += 256) {
    const int hg = lane & 3, kq = warp * 8 + (lane >> 2);
    const int k = kbase + kq * 4;
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int hq = H / 4;                                  // h range of this lane: [hg hq, (hg + 1) hq)
    if (k < H) {
      for (int hb = hg * hq; hb < (hg + 1) * hq; hb += 16) {
        float4 w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) w[u] = __ldg(reinterpret_cast<const float4*>(a.W1v + (size_t)(hb + u) * H + k));
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nv) {
              const float dh = s_dhv[v * H + hb + u];
              acc[v].x = fmaf(dh, w[u].x, acc[v].x);
              acc[v].y = fmaf(dh, w[u].y, acc[v].y);
              acc[v].z = fmaf(dh, w[u].z, acc[v].z);
              acc[v].w = fmaf(dh, w[u].w, acc[v].w);
            }
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {                           // fold the four h quarters (lanes hg = 0..3 of the group)
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
        acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
        acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
        acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
      }
    }
    if (hg == 0 && k < H) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (v < nv) *reinterpret_cast<float4*>(s_fv + v * H + k) = acc[v];      // s_fv is free after phase 2
    }
  }
  row_sync();
  TAIL_MARK();
  for (int k = tid; k < H; k += kRowThreads) {              // thread per feature: classifier part, dropout, store
    float cls[V];
#pragma unroll
    for (int v = 0; v < V; ++v) cls[v] = 0.f;
    for (int c0 = 0; c0 < C; c0 += 16) {
      float w[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) w[u] = (c0 + u < C) ? __ldg(a.Wc + (size_t)(c0 + u) * H + k) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (c0 + u < C) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nv) cls[v] = fmaf(s_gv[v * C + c0 + u], w[u], cls[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (v < nv) {
        const size_t ge = (size_t)(v0 + v) * H + k;
        const float g = (cls[v] - beta1 * s_fv[v * H + k]) * drop_factor(a.drop_v, ge);
        s_G[v * H + k] = g;
        a.G[ge] = g;
      }
  }
  row_sync();
  TAIL_MARK();
  // ---- phase 8: attention gradient, relation-discriminator hidden gradient, frame-discriminator hidden gradient
  for (int it = warp; it < nv * R; it += 8) {               // relattn_bwd_pre
    const int v = it / R, i = it - v * R;
    float pt0 = s_gr[(v * R + i) * 2], pt1 = s_gr[(v * R + i) * 2 + 1];
    if (a.use_attn) {                                       // attention weights are NOT detached
      const float* fr = s_fr + (v * R + i) * H;
      const float* gr = s_G + v * H;
      float dw = 0.f;
      for (int h = lane; h < H; h += 32) dw = fmaf(gr[h], fr[h], dw);
      dw = warp_sum(dw);
      const Attn2 x = attn_from_logits(s_pr[(v * R + i) * 2], s_pr[(v * R + i) * 2 + 1]);
      pt0 += dw * x.q0 * (x.lq0 + x.ent);
      pt1 += dw * x.q1 * (x.lq1 + x.ent);
    }
    const size_t o = (size_t)(v0 + v) * R + i;
    if (lane == 0) {
      a.Pt[o * 2] = pt0;
      a.Pt[o * 2 + 1] = pt1;
    }
    const float* w0 = args.W2r.p[i];
    const float* hr = a.hid_r + ((size_t)i * M + (v0 + v)) * H;
    float* dh = a.dHid + ((size_t)i * M + (v0 + v)) * H;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 h = __ldcg(reinterpret_cast<const float4*>(hr + k));
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(w0 + k));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(w0 + H + k));
      float4 d;
      d.x = h.x > 0.f ? fmaf(pt0, x0.x, pt1 * x1.x) : 0.f;
      d.y = h.y > 0.f ? fmaf(pt0, x0.y, pt1 * x1.y) : 0.f;
      d.z = h.z > 0.f ? fmaf(pt0, x0.z, pt1 * x1.z) : 0.f;
      d.w = h.w > 0.f ? fmaf(pt0, x0.w, pt1 * x1.w) : 0.f;
      *reinterpret_cast<float4*>(dh + k) = d;
    }
  }
  for (int it = warp; it < nv * T; it += 8) {               // dHf = (g_frame W2f) * 1[hid_f > 0]
    const int v = it / T, t = it - v * T;
    const float g0 = s_gf[(v * T + t) * 2], g1 = s_gf[(v * T + t) * 2 + 1];
    const size_t row = (size_t)(v0 + v) * T + t;
    const float* hr = a.hid_f + row * F;
    float* dh = a.dHf + row * F;
    for (int k = lane * 4; k < F; k += 128) {
      const float4 h = __ldcg(reinterpret_cast<const float4*>(hr + k));
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(a.W2f + k));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(a.W2f + F + k));
      float4 d;
      d.x = h.x > 0.f ? fmaf(g0, x0.x, g1 * x1.x) : 0.f;
      d.y = h.y > 0.f ? fmaf(g0, x0.y, g1 * x1.y) : 0.f;
      d.z = h.z > 0.f ? fmaf(g0, x0.z, g1 * x1.z) : 0.f;
      d.w = h.w > 0.f ? fmaf(g0, x0.w, g1 * x1.w) : 0.f;
      *reinterpret_cast<float4*>(dh + k) = d;
    }
  }
  TAIL_MARK();
#undef TAIL_MARK
}

// ---- column sums as tasks --------------------------------------------------------------------------------
// Same arithmetic and the same fixed summation order as wcolsum_stage1/2 (rowops.cuh); the block / split indices
// arrive as arguments instead of blockIdx.  part: (job, column block cb, row split) -> partial[split]; the job's
// reduce task then sums its splits in order.  tid in [0, 256) = (lane, warp) = (threadIdx.x, threadIdx.y) there.
template <int KMAX, bool VEC>
__device__ __forceinline__ void colsum_part_body(const WColsumJob& j, float4 (*red)[33], const int cb, const int split,
                                                 const int k0, const int lane, const int warp) {
  const int n = cb * 128 + lane * 4;
  const int nsplit = j.nsplit;
  float4 acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = f4_zero();
  auto ldx = [&](const float* X, int r) -> float4 {
    const float* q = X + (size_t)r * j.ld + n;
    if (VEC) return __ldcg(reinterpret_cast<const float4*>(q));
    float4 v = f4_zero();
    if (n < j.N) v.x = __ldcg(q);
    if (n + 1 < j.N) v.y = __ldcg(q + 1);
    if (n + 2 < j.N) v.z = __ldcg(q + 2);
    if (n + 3 < j.N) v.w = __ldcg(q + 3);
    return v;
  };
  if (n < j.N) {
    for (int sg = 0; sg < j.nseg; ++sg) {
      const float* X = j.X[sg];
      const float* P = j.P[sg];
      const int rows = j.rows[sg];
      const int per = (rows + nsplit - 1) / nsplit;
      const int r1 = min(rows, (split + 1) * per);
      int r = split * per + warp;
      for (; r + 24 < r1; r += 32) {
        const float4 x0 = ldx(X, r), x1 = ldx(X, r + 8), x2 = ldx(X, r + 16), x3 = ldx(X, r + 24);
        if (P == nullptr) {
          acc[0].x += (x0.x + x1.x) + (x2.x + x3.x);
          acc[0].y += (x0.y + x1.y) + (x2.y + x3.y);
          acc[0].z += (x0.z + x1.z) + (x2.z + x3.z);
          acc[0].w += (x0.w + x1.w) + (x2.w + x3.w);
        } else {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k0 + k < j.N2) {
              f4_fma(acc[k], __ldcg(P + (size_t)r * j.ldp + k0 + k), x0);
              f4_fma(acc[k], __ldcg(P + (size_t)(r + 8) * j.ldp + k0 + k), x1);
              f4_fma(acc[k], __ldcg(P + (size_t)(r + 16) * j.ldp + k0 + k), x2);
              f4_fma(acc[k], __ldcg(P + (size_t)(r + 24) * j.ldp + k0 + k), x3);
            }
        }
      }
      for (; r < r1; r += 8) {
        const float4 x = ldx(X, r);
        if (P == nullptr) {
          acc[0].x += x.x;
          acc[0].y += x.y;
          acc[0].z += x.z;
          acc[0].w += x.w;
        } else {
#pragma unroll
          for (int k = 0; k < KMAX; ++k)
            if (k0 + k < j.N2) f4_fma(acc[k], __ldcg(P + (size_t)r * j.ldp + k0 + k), x);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k0 + k >= j.N2) break;
    red[warp][lane] = acc[k];
    row_sync();
    if (warp == 0 && n < j.N) {
      float4 t = f4_zero();
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const float4 v = red[y][lane];
        t.x += v.x;
        t.y += v.y;
        t.z += v.z;
        t.w += v.w;
      }
      float* o = j.partial + ((size_t)split * j.N2 + k0 + k) * j.N + n;
      if (VEC) {
        *reinterpret_cast<float4*>(o) = t;
      } else {
        o[0] = t.x;
        if (n + 1 < j.N) o[1] = t.y;
        if (n + 2 < j.N) o[2] = t.z;
        if (n + 3 < j.N) o[3] = t.w;
      }
    }
    row_sync();
  }
}

__device__ __forceinline__ void colsum_part_task(const WColsumJob& job_in, float* __restrict__ sm, const int cb,
                                                 const int split, const int tid) {
  const WColsumJob j = job_in;      // register copy: no generic reloads behind the stores
  float4(*red)[33] = reinterpret_cast<float4(*)[33]>(sm);
  const int lane = tid & 31, warp = tid >> 5;
  if (cb * 128 >= j.N || split >= j.nsplit) return;
  if (j.vec4) {
    if (j.N2 <= 1) {
      colsum_part_body<1, true>(j, red, cb, split, 0, lane, warp);
    } else if (j.N2 <= 2) {
      colsum_part_body<2, true>(j, red, cb, split, 0, lane, warp);
    } else {
      for (int k0 = 0; k0 < j.N2; k0 += 4) colsum_part_body<4, true>(j, red, cb, split, k0, lane, warp);
    }
  } else {
    if (j.N2 <= 1) {
      colsum_part_body<1, false>(j, red, cb, split, 0, lane, warp);
    } else if (j.N2 <= 2) {
      colsum_part_body<2, false>(j, red, cb, split, 0, lane, warp);
    } else {
      for (int k0 = 0; k0 < j.N2; k0 += 4) colsum_part_body<4, false>(j, red, cb, split, k0, lane, warp);
    }
  }
}

// out[k, n] = sum_split partial[split, k, n] in split order (the whole job: N2*N outputs, 256 threads)
__device__ __forceinline__ void colsum_reduce_task(const WColsumJob& job_in, const int tid) {
  const WColsumJob j = job_in;
  const int total = j.N2 * j.N;
  const int nsplit = j.nsplit;
  for (int e = tid; e < total; e += kRowThreads) {
    const float* p = j.partial + e;
    float s = 0.f;
    int sp = 0;
    for (; sp + 8 <= nsplit; sp += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldcg(p + (size_t)(sp + i) * total);
      s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    for (; sp < nsplit; ++sp) s += __ldcg(p + (size_t)sp * total);
    j.out[(size_t)(e / j.N) * j.ldo + (e % j.N)] = s;
  }
}

}  // namespace ta3n
