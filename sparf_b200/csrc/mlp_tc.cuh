// Internal interface of the tcgen05 MLP engine (mlp_tc.cu).
#pragma once
#include "common.cuh"

namespace sparf {

// true iff the MLP has the one shape the fused kernels are specialised for (the reference default:
// 8x256 trunk, skip at 4, L_xyz=10, L_view=4, 128-wide colour head)
bool tc_supports(const SparfMLP* mlp);
bool tc_backward_available();
size_t tc_workspace_bytes(const SparfMLP* mlp, int R, int S, int backward, int engine);
int tc_mlp_forward(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                   const float* t, const float* noise, float* sigma, float* rgb, void* workspace,
                   size_t workspace_bytes, cudaStream_t st);
int tc_mlp_backward(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                    const float* t, const float* noise, const float* d_sigma, const float* d_rgb,
                    const SparfMLPGrad* grad, float* d_origins, float* d_dirs, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);

// tape API: the training forward dumps what the backward needs (no recompute); 0 bytes = not available
size_t tc_tape_bytes(const SparfMLP* mlp, int R, int S);
int tc_mlp_forward_tape(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                        const float* t, const float* noise, float* sigma, float* rgb, void* tape, size_t tape_bytes,
                        void* workspace, size_t workspace_bytes, cudaStream_t st);
int tc_mlp_backward_tape(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                         const float* t, const float* sigma, const float* rgb, const float* d_sigma, const float* d_rgb,
                         const SparfMLPGrad* grad, float* d_origins, float* d_dirs, void* tape, size_t tape_bytes,
                         void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace sparf
