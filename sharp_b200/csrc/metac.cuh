// metac.cuh -- argument blocks of the wMetaC / sMetaC kernels (internal).
#pragma once
#include "internal.cuh"

namespace sharp {

constexpr int WM_MAXK = 32;   // ensemble members (columns of nC)
constexpr int WM_MAXL = 256;  // largest label code per column
constexpr int WM_MAXU = 512;  // clusters per block after voting

struct WmArgs {
    const int32_t *labels;  // [K][ncells] label codes 1..WM_MAXL (enrp)
    int64_t ncells;
    int K;
    const int64_t *start;   // [T+1] block boundaries (device)
    int capC;               // capacity of clusters per block (allC <= capC)
    int capU;               // capacity of unique(finalC) per block
    int *gid;               // [K][N_t] per block, stored at K*start[t]: global cluster id of (cell, member)
    int *members;           // same shape: cells grouped by cluster, ascending inside
    int *moff;              // [T][capC+1] offsets into the block's members region
    int *col_of;            // [T][capC] member (column) of every cluster
    int *allc;              // [T] number of clusters (negative: over capacity)
    double *w1;             // [ncells]
    double *S, *D, *Dw;     // [T][capC*capC]
    HcProb *probs;          // [T]
    int *ia, *ib;           // [T][capC]
    double *crit;           // [T][capC]
    int *finalc;            // [ncells] meta-cluster id chosen by the vote
    int *fcode;             // [ncells] index of finalc in unique(finalC) of the block
    int *ucount;            // [T]
    int *ulist;             // [T][capU]
    int *status;            // [T]
};

struct SmArgs {
    const int *nc_ptr;      // number of clusters (device)
    const int *status_in;
    HcProb *prob;
    int ld;                 // capacity / leading dimension of S, D
    double *S, *D, *Dw;
    int *ia, *ib;
    double *crit;
    HcParamsDev prm;        // as given by the caller
    int64_t ncells_total;
    HcParamsDev *prm_out;   // after the k-range tweak (device)
    int *tf;                // [ld]
    int *status_out;
};

int launch_wmetac_front(sharp_ctx *c, const WmArgs &A, int T, int max_block_n);
int launch_wmetac_vote(sharp_ctx *c, const WmArgs &A, const SweepOut *outs, int T);
int launch_wmetac_x0(sharp_ctx *c, const WmArgs &A, const SweepOut *outs, int T, int max_block_n, const int *coloff,
                     const int *colmap, const int64_t *out_row, double *x0, int ncol);
int launch_sm_codes(sharp_ctx *c, const WmArgs &A, int T, int *coloff, int *nc_out, int *status_out, int *code,
                    int *corder, int *coff);
int launch_sm_centroids(sharp_ctx *c, const double *E1, int p, const int *corder, const int *coff, const int *nc_ptr,
                        int nc_cap, double *cen, int64_t *counts);
int launch_sm_similarity(sharp_ctx *c, const SmArgs &A, const double *cen, int p, int nc_cap, double *mean, double *sdev);
int launch_sm_finish(sharp_ctx *c, const SmArgs &A, const SweepOut *out);
int launch_sm_relabel(sharp_ctx *c, int64_t ncells, const int *code, const int *tf, int add, const int64_t *out_row, int *out);

}  // namespace sharp
