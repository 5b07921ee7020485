// Cluster-persistent greedy decode kernel of the bf16 tier (decode_mega.cu): host-side argument block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int MEGA_MAX_LAYERS = 8;
constexpr int MEGA_ROWS_PER_GROUP = 80;     // at most this many sequences per cluster
constexpr int MEGA_CLUSTER = 16;            // CTAs per cluster

struct MegaLayerW {                          // bf16 [N][K] K-major; GLU / GeGLU weights row-interleaved (value 2j, gate 2j+1)
    const void* wqkv;                        // [1536][256]  q | k | v
    const void* wo_s; const float* bo_s;     // [512][512], [512]   self-attention output projection (GLU)
    const void* wq_c;                        // [512][256]          cross-attention query projection
    const void* wo_c; const float* bo_c;     // [512][512], [512]
    const void* w1; const float* b1;         // [2048][256], [2048] GeGLU
    const void* w2; const float* b2;         // [256][1024], [256]
};

struct MegaArgs {
    int B, L, V, tcap, eos, nsteps;
    int G;                                   // groups (clusters): group g owns rows [g*B/G, (g+1)*B/G), at most 80
    MegaLayerW layer[MEGA_MAX_LAYERS];
    const void* w_logits; const float* b_logits;           // [V][256], [V]
    const float* tok_emb; const float* pos_emb;            // [V][256], [max_length][256]
    const float* ln_g; const float* ln_b;                  // the stack's shared LayerNorm
    const float* fin_g; const float* fin_b;                // final norm
    // generate-loop state (one step counter / done flag per group of 64 rows)
    int64_t* cur_tok; int* step; int* done_step; int* seen; int64_t* out_ids;
    // workspaces [B][...]
    float* x; float* s; void* qkv; void* o; void* hid; float* part_val; int* part_idx;
    // K/V: self cache [L][B][8][tcap][K64|V64]; encoder memory [L][8][ntok][K64|V64]; enc_off [B+1] token offsets
    void* kv_self; const void* kv_cross; long ntok; const int* enc_off;
    unsigned long long* dbg_time;     // nullable: [16] nanoseconds per phase type summed over CTAs (thread 0 of each), see decode_mega.cu
};

// dims: d_model 256, 8 heads x 64, GeGLU 2048 -> 1024, V <= 1024, L <= MEGA_MAX_LAYERS, 16-CTA clusters schedulable
bool decode_mega_supported(int B, int L, int V, int* why);
cudaError_t launch_decode_mega(const MegaArgs& a, cudaStream_t st);
int decode_mega_groups(int B);          // how a batch of B is cut into groups on this device
int decode_mega_active_clusters();     // 16-CTA clusters of the kernel the device can hold at once (0 = cannot launch)
