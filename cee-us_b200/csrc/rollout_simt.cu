// fp32 (CUDA-core FMA) rollout kernels: the whole horizon of predict_n_steps in ONE launch.
//
//   GNN:  GNNForwardEnsembleModel.predict_n_steps  (mbrl/models/gnn_ensemble.py:876-983) around
//         GraphNeuralNetworkEnsemble.forward       (mbrl/models/torch_parallel_ensembles/gnn_modules.py:173-278)
//   MLP:  ParallelNNDeterministicEnsemble.predict_n_steps (mbrl/models/mlp_ensemble.py:502-611)
//
// This is the CEEUS_PREC_FP32 path (parity bar 1e-5 against the CPU oracle).  A CTA owns one ensemble member and
// one group of trajectories for all H steps: the object states stay in shared memory, every linear layer is a
// 64-row x (16*NC)-column register-tiled SGEMM whose weight rows stream through a cp.async double buffer, and
// bias + LayerNorm + activation are fused into the GEMM epilogue.  The only HBM traffic is the action read and
// one coalesced store of the predicted observation per step.
#include "common.cuh"
#include "cost_device.cuh"
#include "kernels.cuh"

namespace ceeus {

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case CEEUS_ACT_RELU: return fmaxf(x, 0.f);
    case CEEUS_ACT_SILU: return x / (1.f + expf(-x));
    case CEEUS_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// acc[4][NC] (+)= X[64][K] * W[K][16*NC].  Thread (tx = tid&15, ty = tid>>4) owns rows ty*4..+3 and the column
// quads {j*64 + tx*4 .. +3}.  W rows are staged kKChunk at a time through a 2-deep cp.async ring.
// weight rows staged per cp.async chunk: 32, or 16 for 256-wide outputs (keeps the ring at 32 KB so that a hidden-256 GNN fits)
__host__ __device__ constexpr int kchunk_of(int nout) { return nout >= 256 ? 16 : kKChunk; }
__host__ __device__ constexpr int wbuf_floats(int hd) { return 2 * kchunk_of(hd) * hd; }

template <int NC>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ Wg, int Kpad, const float* __restrict__ X, int XS,
                                          float* __restrict__ Wbuf, float (&acc)[4][NC]) {
  constexpr int NOUT = 16 * NC;
  constexpr int kKChunk = kchunk_of(NOUT);  // (shadows the namespace constant inside this function)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;

  const int nchunks = (Kpad + kKChunk - 1) / kKChunk;
  auto load_chunk = [&](int c) {
    const int k0 = c * kKChunk;
    const int rows = min(kKChunk, Kpad - k0);
    float* dst = Wbuf + (c & 1) * kKChunk * NOUT;
    const float* src = Wg + (size_t)k0 * NOUT;
    for (int idx = tid; idx < rows * (NOUT / 4); idx += kSimtThreads) cp_async16(dst + idx * 4, src + idx * 4);
    cp_async_commit();
  };
  load_chunk(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      load_chunk(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int k0 = c * kKChunk;
    const int rows = min(kKChunk, Kpad - k0);
    const float* Wb = Wbuf + (c & 1) * kKChunk * NOUT + tx * 4;
    const float* xr = X + (ty * 4) * XS + k0;
#pragma unroll 2
    for (int k = 0; k < rows; k += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(xr + i * XS + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float* wr = Wb + (k + kk) * NOUT;
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(wr + j * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
            acc[i][j * 4 + 0] = fmaf(av, w.x, acc[i][j * 4 + 0]);
            acc[i][j * 4 + 1] = fmaf(av, w.y, acc[i][j * 4 + 1]);
            acc[i][j * 4 + 2] = fmaf(av, w.z, acc[i][j * 4 + 2]);
            acc[i][j * 4 + 3] = fmaf(av, w.w, acc[i][j * 4 + 3]);
          }
        }
      }
    }
    __syncthreads();  // ring slot (c&1) may be refilled by the prefetch of chunk c+2
  }
}

__device__ __forceinline__ float halfwarp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

// bias + (LayerNorm) + activation, written back IN PLACE over the input tile (models/utils.py:116-126, simple.py:36-50)
template <int NC>
__device__ __forceinline__ void hidden_epilogue(float (&acc)[4][NC], const float* __restrict__ wm, const LayerDesc& ld,
                                                bool ln, int act, float* __restrict__ X, int XS) {
  constexpr int NOUT = 16 * NC;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(wm + ld.b_off + j * 64 + tx * 4));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][j * 4 + 0] += b.x;
      acc[i][j * 4 + 1] += b.y;
      acc[i][j * 4 + 2] += b.z;
      acc[i][j * 4 + 3] += b.w;
    }
  }
  if (ln) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) s += acc[i][c];
      const float mean = halfwarp_sum(s) * (1.f / NOUT);
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float d = acc[i][c] - mean;
        acc[i][c] = d;
        v = fmaf(d, d, v);
      }
      const float rstd = 1.f / sqrtf(halfwarp_sum(v) * (1.f / NOUT) + 1e-5f);
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[i][c] *= rstd;
    }
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(wm + ld.g_off + j * 64 + tx * 4));
      const float4 be = __ldg(reinterpret_cast<const float4*>(wm + ld.be_off + j * 64 + tx * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][j * 4 + 0] = fmaf(acc[i][j * 4 + 0], g.x, be.x);
        acc[i][j * 4 + 1] = fmaf(acc[i][j * 4 + 1], g.y, be.y);
        acc[i][j * 4 + 2] = fmaf(acc[i][j * 4 + 2], g.z, be.z);
        acc[i][j * 4 + 3] = fmaf(acc[i][j * 4 + 3], g.w, be.w);
      }
    }
  }
  // (gemm_tile ended with a __syncthreads: nobody reads X any more)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      float4 o;
      o.x = apply_act(acc[i][j * 4 + 0], act);
      o.y = apply_act(acc[i][j * 4 + 1], act);
      o.z = apply_act(acc[i][j * 4 + 2], act);
      o.w = apply_act(acc[i][j * 4 + 3], act);
      *reinterpret_cast<float4*>(X + (ty * 4 + i) * XS + j * 64 + tx * 4) = o;
    }
  __syncthreads();
}

// output layer: bias only, first N columns of the first `rows` rows go to dst[row*ds + col]
template <int NC>
__device__ __forceinline__ void out_epilogue(float (&acc)[4][NC], const float* __restrict__ wm, const LayerDesc& ld,
                                             float* __restrict__ dst, int ds, int rows) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = j * 64 + tx * 4 + c;
          if (col < ld.N) dst[r * ds + col] = acc[i][j * 4 + c] + __ldg(wm + ld.b_off + col);
        }
    }
  }
  __syncthreads();
}

// Runs the hidden layers of one MLP in place on X (X holds the layer-0 input on entry, the last hidden
// activation on exit).  NCH = hidden width / 16.
template <int NCH>
__device__ __forceinline__ void mlp_hidden(const MlpDesc& m, const float* __restrict__ wm, int act, bool ln,
                                           float* __restrict__ X, int XS, float* __restrict__ Wbuf) {
  float acc[4][NCH];
  for (int l = 0; l < m.n_layers - 1; ++l) {
    const LayerDesc& ld = m.layer[l];
    gemm_tile<NCH>(wm + ld.w_off, ld.Kpad, X, XS, Wbuf, acc);
    hidden_epilogue<NCH>(acc, wm, ld, ln, act, X, XS);
  }
}

template <int NCH>
__device__ __forceinline__ void zero_cols(float* X, int XS, int c0, int c1) {
  const int w = c1 - c0;
  for (int idx = threadIdx.x; idx < kTileRows * w; idx += kSimtThreads) X[(idx / w) * XS + c0 + idx % w] = 0.f;
}

// Tail of a planner launch: this member's rollout of the CTA's trajectories is complete; warp w arrives for the trajectories
// w, w + 8, ... and reduces those whose last ensemble member this was (cost_device.cuh).
__device__ __forceinline__ void simt_fused_tail(const costdev::FusedCost& fc, const RolloutArgs& ra, const ModelDesc& md, int p0, int Gc,
                                                float* scratch, int scratch_floats) {
  __threadfence();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = scratch_floats / (kSimtThreads / 32);
  float* envbuf = per_warp >= md.E * ra.H ? scratch + warp * per_warp : nullptr;  // a slice of the (now idle) GEMM input tile
  for (int b = warp; b < Gc; b += kSimtThreads / 32) {
    const int p = p0 + b;
    const float* srow = ra.start_obs + (size_t)(p / ra.traj_per_start) * md.O;
    const float* grow = ra.goal + (size_t)(p / ra.traj_per_goal) * md.G;
    if (md.E <= 5) costdev::fused_cost_arrive<5>(fc, p, lane, srow, grow, envbuf);
    else costdev::fused_cost_arrive<kMaxEnsemble>(fc, p, lane, srow, grow, envbuf);
  }
  __syncthreads();  // the tile is reused by the next group
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// GNN rollout.  grid = (groups, E); a group is G trajectories with G*N <= 64 node rows; edge tiles hold TPT
// whole trajectories (TPT*N*(N-1) <= 64 rows).
// ------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kSimtThreads, 1) rollout_gnn_simt_kernel(const ModelDesc md, const RolloutArgs ra,
                                                                            const GnnTiling tl, const costdev::FusedCost fc) {
  const bool fused = fc.pred != nullptr;  // planner launch: predictions -> compact scratch, costs reduced in the tail
  constexpr int HD = 16 * NCH;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int e = blockIdx.y;
  const int N = md.N, A = md.A, D = md.D, S = md.S, U = md.U, sd = md.sd, O = md.O;
  const int NA = D + S;          // node attribute dim
  const int NE = N * (N - 1);    // edges per trajectory
  const int G = tl.G, TPT = tl.TPT, XSe = tl.XSe, XSn = tl.XSn;
  const float* __restrict__ wm = ra.weights + (size_t)e * md.member_stride;
  const float* __restrict__ nrm = ra.norm;

  float* Wbuf = smem;                          // weight ring (wbuf_floats)
  float* Xe = Wbuf + wbuf_floats(HD);          // 64 * XSe   edge tile, later the global-MLP tile
  float* Xn = Xe + kTileRows * XSe;            // 64 * XSn   node-MLP tile [node_i | agent | action | agg_i]
  float* aggG = Xn + kTileRows * XSn;          // G * HD     per-trajectory mean of all edge messages
  float* sraw = aggG + G * HD;                 // G * sd     raw (un-normalised) state
  float* snrm = sraw + G * sd;                 // G * (sd+U) normalised [agent | nodes (dyn,stat) | action]
  float* dn = snrm + G * (sd + U);             // 64 * D     node deltas (normalised)
  float* dg = dn + kTileRows * D;              // G * A      agent deltas (normalised)
  float* aggN = dg + G * A;                    // G * HD     ignore_agent_object=false: per-trajectory aggregate of edge_mlp_global

  const float inv_node = md.aggr_mean ? 1.f / fmaxf((float)(N - 1), 1.f) : 1.f;
  const float inv_glob = md.aggr_mean ? 1.f / fmaxf((float)NE, 1.f) : 1.f;
  const bool ln = md.layer_norm != 0;
  const int snw = sd + U;

  for (int grp = blockIdx.x; grp * G < ra.n; grp += gridDim.x) {
    const int p0 = grp * G;
    const int Gc = min(G, ra.n - p0);
    __syncthreads();
    // ---- load start observation, emit traj[0] ----
    for (int idx = tid; idx < G * sd; idx += kSimtThreads) {
      const int b = idx / sd, c = idx % sd;
      float v = 0.f;
      if (b < Gc) v = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
      sraw[idx] = v;
    }
    for (int idx = tid; idx < (fused ? 0 : Gc * O); idx += kSimtThreads) {
      const int b = idx / O, c = idx % O;
      ra.traj[((size_t)e * ra.n + p0 + b) * O + c] = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
    }
    __syncthreads();

    for (int h = 0; h < ra.H; ++h) {
      // ---- normalise state + action (gnn_ensemble.py:190-243) ----
      for (int idx = tid; idx < G * snw; idx += kSimtThreads) {
        const int b = idx / snw, c = idx % snw;
        float v;
        if (c < A) {
          v = (sraw[b * sd + c] - __ldg(nrm + md.nrm.in_agent + c)) / __ldg(nrm + md.nrm.in_agent + A + c);
        } else if (c < sd) {
          // normalised layout: node i at A + i*NA: [dyn(D) | stat(S)]; raw layout: [agent | N*D dyn | N*S stat]
          const int q = c - A, i = q / NA, f = q % NA;
          if (f < D)
            v = (sraw[b * sd + A + i * D + f] - __ldg(nrm + md.nrm.in_dyn + f)) / __ldg(nrm + md.nrm.in_dyn + D + f);
          else
            v = (sraw[b * sd + A + N * D + i * S + (f - D)] - __ldg(nrm + md.nrm.in_stat + (f - D))) /
                __ldg(nrm + md.nrm.in_stat + S + (f - D));
        } else {
          const int u = c - sd;
          float a = 0.f;
          if (b < Gc) a = __ldg(ra.actions + ((size_t)(p0 + b) * ra.H + h) * U + u);
          v = (a - __ldg(nrm + md.nrm.in_action + u)) / __ldg(nrm + md.nrm.in_action + U + u);
        }
        snrm[idx] = v;
      }
      __syncthreads();
      // ---- static part of the node tile: [node_i | agent | action]  (gnn_modules.py:213-223) ----
      {
        const int w = NA + A + U;
        for (int idx = tid; idx < kTileRows * w; idx += kSimtThreads) {
          const int r = idx / w, c = idx % w;
          const int b = r / N, i = r % N;
          float v = 0.f;
          if (r < G * N) {
            if (c < NA) v = snrm[b * snw + A + i * NA + c];
            else if (c < NA + A) v = snrm[b * snw + (c - NA)];
            else v = snrm[b * snw + sd + (c - NA - A)];
          }
          Xn[r * XSn + c] = v;
        }
        // zero the K padding of the node tile (cols beyond w + HD up to Kpad)
        const int kp = md.mlp[1].layer[0].Kpad;
        for (int idx = tid; idx < kTileRows * (kp - w - HD); idx += kSimtThreads)
          Xn[(idx / (kp - w - HD)) * XSn + w + HD + idx % (kp - w - HD)] = 0.f;
      }
      // ---- edge stage ----
      for (int b0 = 0; b0 < G; b0 += TPT) {
        const int e_in = 2 * NA + A + U;
        const int kp = md.mlp[0].layer[0].Kpad;
        for (int idx = tid; idx < kTileRows * kp; idx += kSimtThreads) {
          const int r = idx / kp, c = idx % kp;
          const int bb = r / NE, q = r % NE;
          const int b = b0 + bb;
          float v = 0.f;
          if (bb < TPT && b < G && c < e_in) {
            const int i = q / (N - 1), jj = q % (N - 1);
            const int j = jj + (jj >= i ? 1 : 0);  // row-major (i asc, j asc, j != i): gnn_modules.py:148-170
            if (c < NA) v = snrm[b * snw + A + i * NA + c];
            else if (c < 2 * NA) v = snrm[b * snw + A + j * NA + (c - NA)];
            else if (c < 2 * NA + A) v = snrm[b * snw + (c - 2 * NA)];
            else v = snrm[b * snw + sd + (c - 2 * NA - A)];
          }
          Xe[r * XSe + c] = v;
        }
        __syncthreads();
        mlp_hidden<NCH>(md.mlp[0], wm, md.act, ln, Xe, XSe, Wbuf);
        {
          float acc[4][NCH];
          const LayerDesc& ld = md.mlp[0].layer[md.mlp[0].n_layers - 1];
          gemm_tile<NCH>(wm + ld.w_off, ld.Kpad, Xe, XSe, Wbuf, acc);
          hidden_epilogue<NCH>(acc, wm, ld, false, CEEUS_ACT_NONE, Xe, XSe);  // messages m_ij in place
        }
        // ---- aggregation (models/utils.py:42-51): per source node over its N-1 edges, per sample over all ----
        const int nb = min(TPT, G - b0);
        for (int idx = tid; idx < nb * (N + 1) * HD; idx += kSimtThreads) {
          const int c = idx % HD;
          const int t = idx / HD;
          const int bb = t / (N + 1), i = t % (N + 1);
          const float* src = Xe + (bb * NE) * XSe + c;
          float s = 0.f;
          if (i < N) {
            for (int jj = 0; jj < N - 1; ++jj) s += src[(i * (N - 1) + jj) * XSe];
            Xn[((b0 + bb) * N + i) * XSn + NA + A + U + c] = s * inv_node;
          } else {
            for (int q = 0; q < NE; ++q) s += src[q * XSe];
            aggG[(b0 + bb) * HD + c] = s * inv_glob;
          }
        }
        __syncthreads();
      }
      // ---- node stage ----
      mlp_hidden<NCH>(md.mlp[1], wm, md.act, ln, Xn, XSn, Wbuf);
      {
        float acc[4][4];
        const LayerDesc& ld = md.mlp[1].layer[md.mlp[1].n_layers - 1];
        gemm_tile<4>(wm + ld.w_off, ld.Kpad, Xn, XSn, Wbuf, acc);
        out_epilogue<4>(acc, wm, ld, dn, D, G * N);
      }
      // ---- ignore_agent_object=false: edge_mlp_global on every node's [node_i | agent | action], aggregated per trajectory over
      // its N nodes (gnn_modules.py:98-106, 225-238) ----
      if (md.edge_global) {
        const int w = NA + A + U;
        const int kp = md.mlp[3].layer[0].Kpad;
        for (int idx = tid; idx < kTileRows * kp; idx += kSimtThreads) {
          const int r = idx / kp, c = idx % kp;
          const int b = r / N, i = r % N;
          float v = 0.f;
          if (r < G * N && c < w) {
            if (c < NA) v = snrm[b * snw + A + i * NA + c];
            else if (c < NA + A) v = snrm[b * snw + (c - NA)];
            else v = snrm[b * snw + sd + (c - NA - A)];
          }
          Xe[r * XSe + c] = v;
        }
        __syncthreads();
        mlp_hidden<NCH>(md.mlp[3], wm, md.act, ln, Xe, XSe, Wbuf);
        {
          float acc[4][NCH];
          const LayerDesc& ld = md.mlp[3].layer[md.mlp[3].n_layers - 1];
          gemm_tile<NCH>(wm + ld.w_off, ld.Kpad, Xe, XSe, Wbuf, acc);
          hidden_epilogue<NCH>(acc, wm, ld, false, CEEUS_ACT_NONE, Xe, XSe);
        }
        const float inv_n = md.aggr_mean ? 1.f / (float)N : 1.f;
        for (int idx = tid; idx < G * HD; idx += kSimtThreads) {
          const int b = idx / HD, c = idx % HD;
          float s = 0.f;
          for (int i = 0; i < N; ++i) s += Xe[(b * N + i) * XSe + c];
          aggN[idx] = s * inv_n;
        }
        __syncthreads();
      }
      // ---- global stage: tile [agent | action | (agg of edge_mlp_global) | agg_all]  (gnn_modules.py:244-254) ----
      {
        const int kp = md.mlp[2].layer[0].Kpad;
        const int o2 = md.edge_global ? HD : 0;
        for (int idx = tid; idx < kTileRows * kp; idx += kSimtThreads) {
          const int r = idx / kp, c = idx % kp;
          float v = 0.f;
          if (r < G) {
            if (c < A) v = snrm[r * snw + c];
            else if (c < A + U) v = snrm[r * snw + sd + (c - A)];
            else if (c < A + U + o2) v = aggN[r * HD + (c - A - U)];
            else if (c < A + U + o2 + HD) v = aggG[r * HD + (c - A - U - o2)];
          }
          Xe[r * XSe + c] = v;
        }
        __syncthreads();
        mlp_hidden<NCH>(md.mlp[2], wm, md.act, ln, Xe, XSe, Wbuf);
        float acc[4][4];
        const LayerDesc& ld = md.mlp[2].layer[md.mlp[2].n_layers - 1];
        gemm_tile<4>(wm + ld.w_off, ld.Kpad, Xe, XSe, Wbuf, acc);
        out_epilogue<4>(acc, wm, ld, dg, A, G);
      }
      // ---- de-normalise, residual update (gnn_ensemble.py:302-334, 935-939), emit traj[h+1] ----
      for (int idx = tid; idx < G * sd; idx += kSimtThreads) {
        const int b = idx / sd, c = idx % sd;
        float v;
        if (c < A) {
          v = fmaf(__ldg(nrm + md.nrm.out_agent + A + c), dg[b * A + c], __ldg(nrm + md.nrm.out_agent + c)) + sraw[idx];
        } else if (c < A + N * D) {
          const int q = c - A, i = q / D, f = q % D;
          v = fmaf(__ldg(nrm + md.nrm.out_dyn + D + f), dn[(b * N + i) * D + f], __ldg(nrm + md.nrm.out_dyn + f)) + sraw[idx];
        } else {
          // static features: normalise -> de-normalise round trip with the input normaliser (:320-324)
          const int q = c - A - N * D, i = q / S, f = q % S;
          v = fmaf(__ldg(nrm + md.nrm.in_stat + S + f), snrm[b * snw + A + i * NA + D + f], __ldg(nrm + md.nrm.in_stat + f));
        }
        sraw[idx] = v;
      }
      __syncthreads();
      if (fused) {
        // planner: only the predicted dims (agent + dynamic object features), into each trajectory's block [H][E][Op]
        for (int idx = tid; idx < Gc * fc.Op; idx += kSimtThreads) {
          const int b = idx / fc.Op, c = idx % fc.Op;
          fc.pred[(size_t)(p0 + b) * fc.ps + (size_t)h * fc.hs + (size_t)e * fc.es + c] = sraw[b * sd + c];
        }
      } else {
        float* out = ra.traj + (((size_t)(h + 1) * md.E + e) * ra.n + p0) * O;
        for (int idx = tid; idx < Gc * O; idx += kSimtThreads) {
          const int b = idx / O, c = idx % O;
          out[idx] = c < sd ? sraw[b * sd + c] : __ldg(ra.goal + (size_t)((p0 + b) / ra.traj_per_goal) * md.G + (c - sd));
        }
      }
      // (next iteration starts with reads of sraw only; writes to snrm are ordered by the sync above)
    }
    if (fused && fc.done != nullptr) simt_fused_tail(fc, ra, md, p0, Gc, Xe, kTileRows * XSe);
  }
}


// output layer with an arbitrary destination: f(row, col, value) for the first ld.N columns of the first `rows` rows
template <int NC, typename F>
__device__ __forceinline__ void out_epilogue_map(float (&acc)[4][NC], const float* __restrict__ wm, const LayerDesc& ld, int rows, F f) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = j * 64 + tx * 4 + c;
          if (col < ld.N) f(r, col, acc[i][j * 4 + c] + __ldg(wm + ld.b_off + col));
        }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// Homogeneous GNN rollout (HomogeneousGraphNeuralNetworkEnsemble, gnn_modules.py:281-545; yaml gnn_ensemble_homogeneous.yaml,
// agent_as_global_node: false): the agent is node 0 of a fully connected graph of M = N + 1 nodes.  Per model step:
//   features  f_0 = embed_agent(agent), f_{1+i} = embed_obj([dyn_i, stat_i])                      (:351-381, :506-523)
//   n_mp x {  m_ij = edge_mlp([f_i, f_j, action]);  f_i <- node_mlp([f_i, action, aggr_{j != i} m_ij])  }   (:449-497, :525-526)
//   d_agent = output_head_agent(f_0), d_dyn_i = output_head_objdyn(f_{1+i})                          (:528-545)
// grid = (groups, E); a group is G trajectories with G*M <= 64 node rows; edge tiles hold TPT whole trajectories.
// ------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kSimtThreads, 1) rollout_gnn_homog_simt_kernel(const ModelDesc md, const RolloutArgs ra,
                                                                                  const GnnTiling tl, const costdev::FusedCost fc) {
  const bool fused = fc.pred != nullptr;
  constexpr int HD = 16 * NCH;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int e = blockIdx.y;
  const int N = md.N, A = md.A, D = md.D, S = md.S, U = md.U, sd = md.sd, O = md.O, FD = md.emb;
  const int NA = D + S, M = N + 1, ME = M * (M - 1);
  const int G = tl.G, TPT = tl.TPT, XSe = tl.XSe, XSn = tl.XSn;
  const float* __restrict__ wm = ra.weights + (size_t)e * md.member_stride;
  const float* __restrict__ nrm = ra.norm;

  float* Wbuf = smem;                          // weight ring (wbuf_floats)
  float* Xe = Wbuf + wbuf_floats(HD);          // 64 * XSe   edge tile; also the input tile of the embeddings / output heads
  float* Xn = Xe + kTileRows * XSe;            // 64 * XSn   node-MLP tile [f_i | action | agg_i]
  float* Fm = Xn + kTileRows * XSn;            // 64 * FD    node features, row b * M + node
  float* sraw = Fm + kTileRows * FD;           // G * sd
  float* snrm = sraw + G * sd;                 // G * (sd+U)
  float* dn = snrm + G * (sd + U);             // 64 * D
  float* dg = dn + kTileRows * D;              // G * A
  const float inv_node = md.aggr_mean ? 1.f / (float)(M - 1) : 1.f;
  const bool ln = md.layer_norm != 0;
  const int snw = sd + U;

  for (int grp = blockIdx.x; grp * G < ra.n; grp += gridDim.x) {
    const int p0 = grp * G;
    const int Gc = min(G, ra.n - p0);
    __syncthreads();
    for (int idx = tid; idx < G * sd; idx += kSimtThreads) {
      const int b = idx / sd, c = idx % sd;
      sraw[idx] = b < Gc ? __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c) : 0.f;
    }
    for (int idx = tid; idx < (fused ? 0 : Gc * O); idx += kSimtThreads) {
      const int b = idx / O, c = idx % O;
      ra.traj[((size_t)e * ra.n + p0 + b) * O + c] = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
    }
    __syncthreads();

    for (int h = 0; h < ra.H; ++h) {
      // ---- normalise state + action (gnn_ensemble.py:190-243) ----
      for (int idx = tid; idx < G * snw; idx += kSimtThreads) {
        const int b = idx / snw, c = idx % snw;
        float v;
        if (c < A) {
          v = (sraw[b * sd + c] - __ldg(nrm + md.nrm.in_agent + c)) / __ldg(nrm + md.nrm.in_agent + A + c);
        } else if (c < sd) {
          const int q = c - A, i = q / NA, f = q % NA;
          if (f < D)
            v = (sraw[b * sd + A + i * D + f] - __ldg(nrm + md.nrm.in_dyn + f)) / __ldg(nrm + md.nrm.in_dyn + D + f);
          else
            v = (sraw[b * sd + A + N * D + i * S + (f - D)] - __ldg(nrm + md.nrm.in_stat + (f - D))) /
                __ldg(nrm + md.nrm.in_stat + S + (f - D));
        } else {
          const int u = c - sd;
          const float a = b < Gc ? __ldg(ra.actions + ((size_t)(p0 + b) * ra.H + h) * U + u) : 0.f;
          v = (a - __ldg(nrm + md.nrm.in_action + u)) / __ldg(nrm + md.nrm.in_action + U + u);
        }
        snrm[idx] = v;
      }
      __syncthreads();
      // ---- embeddings: f_0 = embed_agent(agent), f_{1+i} = embed_obj(node_i) ----
      {
        const LayerDesc& la = md.mlp[2].layer[0];
        for (int idx = tid; idx < kTileRows * la.Kpad; idx += kSimtThreads) {
          const int r = idx / la.Kpad, c = idx % la.Kpad;
          Xe[r * XSe + c] = (r < G && c < A) ? snrm[r * snw + c] : 0.f;
        }
        __syncthreads();
        float acc[4][4];
        gemm_tile<4>(wm + la.w_off, la.Kpad, Xe, XSe, Wbuf, acc);
        out_epilogue<4>(acc, wm, la, Fm, M * FD, G);  // row b -> feature row b * M
        const LayerDesc& lo = md.mlp[3].layer[0];
        for (int idx = tid; idx < kTileRows * lo.Kpad; idx += kSimtThreads) {
          const int r = idx / lo.Kpad, c = idx % lo.Kpad;
          Xe[r * XSe + c] = (r < G * N && c < NA) ? snrm[(r / N) * snw + A + (r % N) * NA + c] : 0.f;
        }
        __syncthreads();
        gemm_tile<4>(wm + lo.w_off, lo.Kpad, Xe, XSe, Wbuf, acc);
        out_epilogue_map<4>(acc, wm, lo, G * N, [&](int r, int col, float v) { Fm[((r / N) * M + 1 + r % N) * FD + col] = v; });
      }
      // ---- message passing ----
      for (int mp = 0; mp < md.n_mp; ++mp) {
        // static part of the node tile: [f_i | action]; K padding zeroed
        {
          const int w = FD + U;
          const int kp = md.mlp[1].layer[0].Kpad;
          for (int idx = tid; idx < kTileRows * w; idx += kSimtThreads) {
            const int r = idx / w, c = idx % w;
            float v = 0.f;
            if (r < G * M) v = c < FD ? Fm[r * FD + c] : snrm[(r / M) * snw + sd + (c - FD)];
            Xn[r * XSn + c] = v;
          }
          for (int idx = tid; idx < kTileRows * (kp - w - HD); idx += kSimtThreads)
            Xn[(idx / (kp - w - HD)) * XSn + w + HD + idx % (kp - w - HD)] = 0.f;
        }
        for (int b0 = 0; b0 < G; b0 += TPT) {
          const int e_in = 2 * FD + U;
          const int kp = md.mlp[0].layer[0].Kpad;
          for (int idx = tid; idx < kTileRows * kp; idx += kSimtThreads) {
            const int r = idx / kp, c = idx % kp;
            const int bb = r / ME, q = r % ME, b = b0 + bb;
            float v = 0.f;
            if (bb < TPT && b < G && c < e_in) {
              const int i = q / (M - 1), jj = q % (M - 1);
              const int j = jj + (jj >= i ? 1 : 0);  // (i asc, j asc, j != i); row = source node i (gnn_modules.py:424-447)
              if (c < FD) v = Fm[(b * M + i) * FD + c];
              else if (c < 2 * FD) v = Fm[(b * M + j) * FD + (c - FD)];
              else v = snrm[b * snw + sd + (c - 2 * FD)];
            }
            Xe[r * XSe + c] = v;
          }
          __syncthreads();
          mlp_hidden<NCH>(md.mlp[0], wm, md.act, ln, Xe, XSe, Wbuf);
          {
            float acc[4][NCH];
            const LayerDesc& ld = md.mlp[0].layer[md.mlp[0].n_layers - 1];
            gemm_tile<NCH>(wm + ld.w_off, ld.Kpad, Xe, XSe, Wbuf, acc);
            hidden_epilogue<NCH>(acc, wm, ld, false, CEEUS_ACT_NONE, Xe, XSe);
          }
          const int nb = min(TPT, G - b0);
          for (int idx = tid; idx < nb * M * HD; idx += kSimtThreads) {
            const int c = idx % HD, t = idx / HD;
            const int bb = t / M, i = t % M;
            const float* src = Xe + (bb * ME + i * (M - 1)) * XSe + c;
            float s = 0.f;
            for (int jj = 0; jj < M - 1; ++jj) s += src[jj * XSe];
            Xn[((b0 + bb) * M + i) * XSn + FD + U + c] = s * inv_node;
          }
          __syncthreads();
        }
        mlp_hidden<NCH>(md.mlp[1], wm, md.act, ln, Xn, XSn, Wbuf);
        {
          float acc[4][4];
          const LayerDesc& ld = md.mlp[1].layer[md.mlp[1].n_layers - 1];
          gemm_tile<4>(wm + ld.w_off, ld.Kpad, Xn, XSn, Wbuf, acc);
          out_epilogue<4>(acc, wm, ld, Fm, FD, G * M);  // the new node features
        }
      }
      // ---- output heads ----
      {
        const LayerDesc& la = md.mlp[4].layer[0];
        for (int idx = tid; idx < kTileRows * la.Kpad; idx += kSimtThreads) {
          const int r = idx / la.Kpad, c = idx % la.Kpad;
          Xe[r * XSe + c] = (r < G && c < FD) ? Fm[(r * M) * FD + c] : 0.f;
        }
        __syncthreads();
        float acc[4][4];
        gemm_tile<4>(wm + la.w_off, la.Kpad, Xe, XSe, Wbuf, acc);
        out_epilogue<4>(acc, wm, la, dg, A, G);
        const LayerDesc& lo = md.mlp[5].layer[0];
        for (int idx = tid; idx < kTileRows * lo.Kpad; idx += kSimtThreads) {
          const int r = idx / lo.Kpad, c = idx % lo.Kpad;
          Xe[r * XSe + c] = (r < G * N && c < FD) ? Fm[((r / N) * M + 1 + r % N) * FD + c] : 0.f;
        }
        __syncthreads();
        gemm_tile<4>(wm + lo.w_off, lo.Kpad, Xe, XSe, Wbuf, acc);
        out_epilogue<4>(acc, wm, lo, dn, D, G * N);
      }
      // ---- de-normalise, residual update (gnn_ensemble.py:302-334, 935-939), emit ----
      for (int idx = tid; idx < G * sd; idx += kSimtThreads) {
        const int b = idx / sd, c = idx % sd;
        float v;
        if (c < A) {
          v = fmaf(__ldg(nrm + md.nrm.out_agent + A + c), dg[b * A + c], __ldg(nrm + md.nrm.out_agent + c)) + sraw[idx];
        } else if (c < A + N * D) {
          const int q = c - A, i = q / D, f = q % D;
          v = fmaf(__ldg(nrm + md.nrm.out_dyn + D + f), dn[(b * N + i) * D + f], __ldg(nrm + md.nrm.out_dyn + f)) + sraw[idx];
        } else {
          const int q = c - A - N * D, i = q / S, f = q % S;
          v = fmaf(__ldg(nrm + md.nrm.in_stat + S + f), snrm[b * snw + A + i * NA + D + f], __ldg(nrm + md.nrm.in_stat + f));
        }
        sraw[idx] = v;
      }
      __syncthreads();
      if (fused) {
        for (int idx = tid; idx < Gc * fc.Op; idx += kSimtThreads) {
          const int b = idx / fc.Op, c = idx % fc.Op;
          fc.pred[(size_t)(p0 + b) * fc.ps + (size_t)h * fc.hs + (size_t)e * fc.es + c] = sraw[b * sd + c];
        }
      } else {
        float* out = ra.traj + (((size_t)(h + 1) * md.E + e) * ra.n + p0) * O;
        for (int idx = tid; idx < Gc * O; idx += kSimtThreads) {
          const int b = idx / O, c = idx % O;
          out[idx] = c < sd ? sraw[b * sd + c] : __ldg(ra.goal + (size_t)((p0 + b) / ra.traj_per_goal) * md.G + (c - sd));
        }
      }
    }
    if (fused && fc.done != nullptr) simt_fused_tail(fc, ra, md, p0, Gc, Xe, kTileRows * XSe);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Dense-MLP rollout (mlp_ensemble.py:530-573): x = norm([state, action]) -> MLP -> state += denorm(y).
// grid = (groups of 64 trajectories, E)
// ------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kSimtThreads, 1) rollout_mlp_simt_kernel(const ModelDesc md, const RolloutArgs ra,
                                                                            const int XS, const costdev::FusedCost fc) {
  const bool fused = fc.pred != nullptr;
  constexpr int HD = 16 * NCH;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int e = blockIdx.y;
  const int U = md.U, sd = md.sd, O = md.O;
  const float* __restrict__ wm = ra.weights + (size_t)e * md.member_stride;
  const float* __restrict__ nrm = ra.norm;
  float* Wbuf = smem;                   // 2 * kKChunk * HD
  float* X = Wbuf + 2 * kKChunk * HD;   // 64 * XS
  float* sraw = X + kTileRows * XS;     // 64 * sd
  float* dl = sraw + kTileRows * sd;    // 64 * sd   normalised deltas
  const int in = sd + U;
  const MlpDesc& m = md.mlp[0];

  for (int grp = blockIdx.x; grp * kTileRows < ra.n; grp += gridDim.x) {
    const int p0 = grp * kTileRows;
    const int Gc = min(kTileRows, ra.n - p0);
    __syncthreads();
    for (int idx = tid; idx < kTileRows * sd; idx += kSimtThreads) {
      const int b = idx / sd, c = idx % sd;
      sraw[idx] = b < Gc ? __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c) : 0.f;
    }
    for (int idx = tid; idx < (fused ? 0 : Gc * O); idx += kSimtThreads) {
      const int b = idx / O, c = idx % O;
      ra.traj[((size_t)e * ra.n + p0 + b) * O + c] = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
    }
    __syncthreads();
    for (int h = 0; h < ra.H; ++h) {
      const int kp = m.layer[0].Kpad;
      for (int idx = tid; idx < kTileRows * kp; idx += kSimtThreads) {
        const int b = idx / kp, c = idx % kp;
        float v = 0.f;
        if (c < in) {
          float x;
          if (c < sd) x = sraw[b * sd + c];
          else x = b < Gc ? __ldg(ra.actions + ((size_t)(p0 + b) * ra.H + h) * U + (c - sd)) : 0.f;
          v = (x - __ldg(nrm + md.nrm.in_all + c)) / __ldg(nrm + md.nrm.in_all + in + c);
        }
        X[b * XS + c] = v;
      }
      __syncthreads();
      mlp_hidden<NCH>(m, wm, md.act, md.layer_norm != 0, X, XS, Wbuf);
      {
        const LayerDesc& ld = m.layer[m.n_layers - 1];
        // output width sd <= 64 or <= 128
        if (ld.Npad == 64) {
          float acc[4][4];
          gemm_tile<4>(wm + ld.w_off, ld.Kpad, X, XS, Wbuf, acc);
          out_epilogue<4>(acc, wm, ld, dl, sd, kTileRows);
        } else {
          float acc[4][8];
          gemm_tile<8>(wm + ld.w_off, ld.Kpad, X, XS, Wbuf, acc);
          out_epilogue<8>(acc, wm, ld, dl, sd, kTileRows);
        }
      }
      for (int idx = tid; idx < kTileRows * sd; idx += kSimtThreads) {
        const int c = idx % sd;
        sraw[idx] += fmaf(dl[idx], __ldg(nrm + md.nrm.out_all + sd + c), __ldg(nrm + md.nrm.out_all + c));
      }
      __syncthreads();
      if (fused) {
        for (int idx = tid; idx < Gc * sd; idx += kSimtThreads) {
          const int b = idx / sd, c = idx % sd;
          fc.pred[(size_t)(p0 + b) * fc.ps + (size_t)h * fc.hs + (size_t)e * fc.es + c] = sraw[idx];
        }
      } else {
        float* out = ra.traj + (((size_t)(h + 1) * md.E + e) * ra.n + p0) * O;
        for (int idx = tid; idx < Gc * O; idx += kSimtThreads) {
          const int b = idx / O, c = idx % O;
          out[idx] = c < sd ? sraw[b * sd + c] : __ldg(ra.goal + (size_t)((p0 + b) / ra.traj_per_goal) * md.G + (c - sd));
        }
      }
    }
    if (fused && fc.done != nullptr) simt_fused_tail(fc, ra, md, p0, Gc, X, kTileRows * XS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------------
static int pick_xs(int kmax) {  // smallest stride >= kmax with stride % 8 == 4: conflict-free float4 row reads
  int xs = round_up(kmax, 4);
  while (xs % 8 != 4) xs += 4;
  return xs;
}

static GnnTiling make_homog_tiling(const ModelDesc& md) {
  GnnTiling t{};
  const int M = md.N + 1, ME = M * (M - 1), Hd = md.Hd;
  t.TPT = kTileRows / ME;
  if (t.TPT < 1) return t;
  const int gmax = kTileRows / M;
  t.G = (gmax / t.TPT) * t.TPT;
  if (t.G < t.TPT) t.G = t.TPT <= gmax ? t.TPT : 0;
  int kmax_e = md.mlp[0].layer[0].Kpad > Hd ? md.mlp[0].layer[0].Kpad : Hd;
  for (int m = 2; m < 6; ++m)
    if (md.mlp[m].layer[0].Kpad > kmax_e) kmax_e = md.mlp[m].layer[0].Kpad;
  t.XSe = pick_xs(kmax_e);
  const int kn = md.mlp[1].layer[0].Kpad;
  t.XSn = pick_xs(kn > Hd ? kn : Hd);
  const size_t fl = (size_t)wbuf_floats(Hd) + (size_t)kTileRows * t.XSe + (size_t)kTileRows * t.XSn + (size_t)kTileRows * md.emb +
                    (size_t)t.G * md.sd + (size_t)t.G * (md.sd + md.U) + (size_t)kTileRows * md.D + (size_t)t.G * md.A;
  t.smem_bytes = fl * sizeof(float);
  return t;
}

GnnTiling make_gnn_tiling(const ModelDesc& md) {
  GnnTiling t{};
  const int NE = md.N * (md.N - 1);
  t.TPT = kTileRows / NE;
  if (t.TPT < 1) return t;  // unsupported (N >= 9), caller checks
  const int gmax = kTileRows / md.N;
  t.G = (gmax / t.TPT) * t.TPT;
  if (t.G < t.TPT) t.G = t.TPT <= gmax ? t.TPT : 0;
  const int Hd = md.Hd;
  if (md.variant == 1) return make_homog_tiling(md);
  const int ke = md.mlp[0].layer[0].Kpad, kg = md.mlp[2].layer[0].Kpad, kn = md.mlp[1].layer[0].Kpad;
  int kmax_e = ke > Hd ? ke : Hd;
  if (kg > kmax_e) kmax_e = kg;
  if (md.edge_global && md.mlp[3].layer[0].Kpad > kmax_e) kmax_e = md.mlp[3].layer[0].Kpad;
  t.XSe = pick_xs(kmax_e);
  t.XSn = pick_xs(kn > Hd ? kn : Hd);
  const size_t fl = (size_t)wbuf_floats(Hd) + (size_t)kTileRows * t.XSe + (size_t)kTileRows * t.XSn + (size_t)t.G * Hd +
                    (size_t)t.G * md.sd + (size_t)t.G * (md.sd + md.U) + (size_t)kTileRows * md.D + (size_t)t.G * md.A +
                    (md.edge_global ? (size_t)t.G * Hd : 0);
  t.smem_bytes = fl * sizeof(float);
  return t;
}

template <int NCH>
static cudaError_t launch_gnn(const ModelDesc& md, const RolloutArgs& ra, const GnnTiling& tl, int num_sms,
                              const costdev::FusedCost& fc, cudaStream_t st) {
  auto kern = md.variant == 1 ? rollout_gnn_homog_simt_kernel<NCH> : rollout_gnn_simt_kernel<NCH>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem_bytes);
  if (err != cudaSuccess) return err;
  const int groups = (ra.n + tl.G - 1) / tl.G;
  dim3 grid(groups, md.E);
  (void)num_sms;
  kern<<<grid, kSimtThreads, tl.smem_bytes, st>>>(md, ra, tl, fc);
  return cudaGetLastError();
}

cudaError_t launch_rollout_gnn_simt(const ModelDesc& md, const RolloutArgs& ra, int num_sms, const costdev::FusedCost& fc,
                                    cudaStream_t st) {
  const GnnTiling tl = make_gnn_tiling(md);
  if (tl.G < 1 || tl.TPT < 1) return cudaErrorInvalidValue;
  switch (md.Hd) {
    case 64: return launch_gnn<4>(md, ra, tl, num_sms, fc, st);
    case 128: return launch_gnn<8>(md, ra, tl, num_sms, fc, st);
    case 256: return launch_gnn<16>(md, ra, tl, num_sms, fc, st);  // robodesk/common/gnn_ensemble.yaml
    default: return cudaErrorInvalidValue;
  }
}

template <int NCH>
static cudaError_t launch_mlp(const ModelDesc& md, const RolloutArgs& ra, const costdev::FusedCost& fc, cudaStream_t st) {
  auto kern = rollout_mlp_simt_kernel<NCH>;
  const int Hd = md.Hd;
  const int k0 = md.mlp[0].layer[0].Kpad;
  const int XS = pick_xs(k0 > Hd ? k0 : Hd);
  const size_t bytes = ((size_t)2 * kKChunk * Hd + (size_t)kTileRows * XS + (size_t)2 * kTileRows * md.sd) * sizeof(float);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err != cudaSuccess) return err;
  dim3 grid((ra.n + kTileRows - 1) / kTileRows, md.E);
  kern<<<grid, kSimtThreads, bytes, st>>>(md, ra, XS, fc);
  return cudaGetLastError();
}

cudaError_t launch_rollout_mlp_simt(const ModelDesc& md, const RolloutArgs& ra, const costdev::FusedCost& fc, cudaStream_t st) {
  switch (md.Hd) {
    case 64: return launch_mlp<4>(md, ra, fc, st);
    case 128: return launch_mlp<8>(md, ra, fc, st);
    case 256: return launch_mlp<16>(md, ra, fc, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ceeus
