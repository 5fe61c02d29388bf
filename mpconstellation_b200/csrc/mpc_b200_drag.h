// mpc_b200_drag.h -- interface between mpc_b200.cu and mpc_b200_drag.cu (the drag-branch kernels' translation unit).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct MpcDragLaunch {
    const double *x, *u, *tf;
    const void *disc_params;   // mpc::DiscParams
    const void *dst_tab;       // mpc::DstTab
    const void *drag_lin;      // mpc::DragLin
    double kf;                 // drag of the DYNAMICS
    int include_j2, n_sats, K, n_sub;
    int variant, block;        // adaptive: 1 = the round-1 kernel; CTA size 32 or 160
    int drag, ucols;           // variant 1 only: without the drag branch (A/B build of round 1), u on its own grid
    double rtol, atol, max_step;
    long long pitch, offset;
    int32_t *status, *n_nodes;
    cudaStream_t stream;
};

size_t mpc_drag_sizeof(int which);   // 0 DiscParams, 1 DstTab, 2 DragLin: layout check between the two units
cudaError_t mpc_drag_launch_fixed(const MpcDragLaunch *a);
cudaError_t mpc_drag_launch_adaptive(const MpcDragLaunch *a);
