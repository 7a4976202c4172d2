// mega.cuh -- host / device interface of the persistent layer kernel (layer_mega.cu) and the peer-inbox layout it shares with tp.cu.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200 {

// ---- tensor-parallel peer inboxes (tp.cu, layer_mega.cu) ------------------------------------------------------------------------
struct PeerSet { char* p[8]; };

struct TpInboxLayout {
    size_t gather_bytes, bcast_off, epoch_off, total;
    int rows_owned;
    __host__ __device__ TpInboxLayout(int world, int rows_max, int n) {
        rows_owned = (rows_max + world - 1) / world;
        gather_bytes = (size_t)2 * world * rows_owned * n * 8;              // [par][src][owned row][n] x {f32, epoch}
        bcast_off = (gather_bytes + 255) & ~(size_t)255;
        const size_t bcast_bytes = (size_t)2 * rows_max * (n / 2) * 8;       // [par][row][n/2] x {half2, epoch}
        epoch_off = (bcast_off + bcast_bytes + 255) & ~(size_t)255;
        total = epoch_off + (((size_t)rows_max * 4 + 4 + 255) & ~(size_t)255);        // epochs (+ one spare word)
    }
};

// ---- persistent layer kernel ------------------------------------------------------------------------------------------------------
constexpr int kMegaMaxPhases = 4;
constexpr int kMegaMaxMaps = 10;
constexpr int kMegaMaxSlabs = 8;              // most CTAs that may share one 128-row tile (wo / w2 / QKV phases); 4 for gate | up

enum MegaEop { kEopNone = 0, kEopNorm = 1, kEopSilu = 2, kEopTpNorm = 3 };

struct MegaPhase {
    // GEMM: y[sg] (slab 0) [m][ldy] f32, slab s at + s * slab_stride; up to three weight matrices on one activation
    int x_map, w_map[3];
    int n[3], tile_end[3];
    float* y[3];
    long long ldy, slab_stride;
    int nsb, n_tiles;                 // n_tiles = 0: no GEMM in this phase (only the elementwise op)
    // elementwise op that turns the PREVIOUS phase's slabs into this phase's activations (kEopNone: they exist before the launch)
    int eop;
    const float* norm_w;              // kEopNorm / kEopTpNorm
    void* act_out;                    // f16 K4 [m][k]: what x_map reads
};

struct MegaParams {
    CUtensorMap maps[kMegaMaxMaps];
    MegaPhase phase[kMegaMaxPhases];
    int n_phases, m, hidden;
    float eps;
    float* x;                         // residual stream f32 [m][hidden]
    uint32_t* counters;               // 2 * kMegaMaxPhases words, zero before the launch
    uint32_t* error_word;             // host-mapped: 2 = a grid-wide wait gave up
    // tensor parallel (kEopTpNorm)
    PeerSet peers;
    int tp_rank, tp_world, rows_max;
    unsigned long long tp_timeout_ns;
    uint32_t* timeout_word;
    int early_trigger;                // 1: griddepcontrol.launch_dependents right after the prologue (B200_MEGA_TRIGGER; default 0: measured 9 % slower)
    long long* trace;                 // profiling aid (B200_MEGA_TRACE): [CTA][phase][8] clock64 stamps, see layer_mega.cu
};

int mega_grid();
int mega_phase_slabs(int n_tiles, int nsb, int G);
bool mega_supported(int m, int hidden, int k_max);
bool mega_make_w_map(CUtensorMap* map, const void* w, int n, int k);
bool mega_make_x_map(CUtensorMap* map, const void* x_f16, int m, int k);
void mega_launch(const MegaParams& P, cudaStream_t st);
void mega_eop_launch(const MegaParams& P, int eop, cudaStream_t st);

#ifdef __CUDACC__
// ---- deterministic split-K bookkeeping (the same arithmetic on the producer and the consumer side) --------------------------
// CTA b owns units [total * b / G, total * (b + 1) / G) of the flattened (tile, super-block) space, so the owner of unit u is
// ((u + 1) * G - 1) / total.  A tile's partial sums live in slabs 0 .. count-1 in the order of the CTAs that computed them.
__device__ __forceinline__ uint32_t unit_owner(uint32_t u, uint32_t total, uint32_t G) { return ((u + 1u) * G - 1u) / total; }
__device__ __forceinline__ int tile_slabs(uint32_t tile, uint32_t nsb, uint32_t total, uint32_t G) {
    return (int)(unit_owner((tile + 1u) * nsb - 1u, total, G) - unit_owner(tile * nsb, total, G)) + 1;
}
#endif

// how a consumer outside the kernel (RoPE + cache write after the QKV phase) finds the partial sums of a tile
struct SlabInfo {
    long long slab_stride;            // 0: single buffer (legacy accumulate path)
    int nsb, n_tiles, grid;
    int seg_tile0[3];                 // first tile of the q / k / v segment
};

}  // namespace b200
