// Shared device/host declarations of libceeus_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ceeus_b200.h"

namespace ceeus {

constexpr int kMaxLayers = 4;    // linear layers per MLP (hidden layers + output layer)
constexpr int kMaxEnsemble = 8;  // E
constexpr int kTileRows = 64;    // rows of one SIMT GEMM tile
constexpr int kSimtThreads = 256;
constexpr int kKChunk = 32;      // weight rows staged per cp.async chunk

// One linear layer of one ensemble member inside the packed weight buffer (offsets in floats from the
// member base).  W is row-major [Kpad][Npad] (the reference's x @ W orientation), zero padded.
struct LayerDesc {
  int K, Kpad, N, Npad;
  int w_off, b_off, g_off, be_off;  // weight, bias, LN gamma, LN beta (g/be unused for the output layer)
};

struct MlpDesc {
  int n_layers;  // linear layers = hidden layers + 1
  LayerDesc layer[kMaxLayers];
};

// Affine normalisers: x_n = (x - mean) / std ; x = std * x_n + mean   (torch_helpers.py:783-791, 840-868)
struct NormDesc {
  // offsets (floats) into the normaliser buffer: mean then std per group
  int in_agent, in_dyn, in_stat, in_action, out_agent, out_dyn;  // GNN
  int in_all, out_all;                                           // dense MLP
};

struct ModelDesc {
  int kind;  // CEEUS_MODEL_*
  int E, N, A, D, S, U, G, Hd, L;
  int act, layer_norm, aggr_mean;
  int sd;  // state dim = A + N (D+S)
  int O;   // obs dim = sd + G
  long long member_stride;  // floats per member in the packed weight buffer
  // GNN, agent as global node: edge, node, global (, edge_mlp_global); homogeneous GNN: edge, node, embed_agent, embed_obj,
  // output_head_agent, output_head_objdyn; dense MLP: [0]
  MlpDesc mlp[6];
  int variant;              // CeeusConfig::gnn_variant
  int edge_global;          // ignore_agent_object: false
  int n_mp;                 // message-passing rounds (homogeneous variant)
  int emb;                  // embedding_dim (homogeneous variant)
  NormDesc nrm;
};

struct CostDesc {
  int curiosity, extrinsic, cost_along;  // CEEUS_COST_*
  int env_kind, cost_case, sparse, shaped;
  int use_std;
  float extrinsic_scale, threshold, buffer_threshold;
  float coord_weights[3], buffer_xyz[3], cost_thres[3], gripper_init[3];  // Flip / Slide
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace ceeus
