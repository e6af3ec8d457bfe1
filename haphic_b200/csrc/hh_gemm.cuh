// Internal interface of the tensor-core pre-expansion (hh_gemm.cu), used by hh_mcl.cu.
#pragma once
#include "hh_common.cuh"
#include "hh_internal.cuh"
#include <vector>

enum { HH_GEMM_DIRECT = 1, HH_GEMM_MIRROR = 2 };

// one output tile of S = C . M0^T: tile coordinates in units of the tile size (hh_gemm_tile_size()), up to two ranges
// of 64-wide k-blocks that are accumulated, and where the result goes
struct hh_gemm_item {
    int m_tile, n_tile;
    int kb_lo[2], kb_hi[2];
    int flags;
    int pad;
};

struct hh_gemm_stats {
    int a_planes, passes, cta_group, stages, chunk_kb;
    int clipped;           // 1: counts above 256 were clipped and the caller owes the sparse correction
    float densify_ms, gemm_ms;
    double flops;          // tensor-core flops issued (2 * M * N * K * passes over all tiles)
};

int hh_gemm_tile_size();
int hh_gemm_items_full(int n, int col_lo, int col_hi, std::vector<hh_gemm_item>& out);
// M1[:, col_lo:col_hi] (dense column-major, leading dimension ld, zero-filled first) from the raw link matrix
int hh_gemm_preexpand(hh_ctx* ctx, const hh_matrix* m, int col_lo, int col_hi, float* d_m1, long long ld, const hh_gemm_item* h_items,
                      int n_items, hh_gemm_stats* st);
