// Cross-translation-unit kernel launchers.
#pragma once
namespace b200rl {
// wgrad_tc.cu: dW_ih, dW_hh, db_ih, db_hh partial slabs on tcgen05 (3xTF32 split)
int launch_wgrad_tc(const float* dpre, const float* h, const float* hx0, const float* u,
                    const float* Wt, const float* bt, float* partials, int nparts, long long P,
                    long long off_Wih, long long off_Whh, long long off_bih, long long off_bhh,
                    int B, int T, int Fin, void* stream);
// dx_tc.cu: torso gradients (dx GEMM on tcgen05 + ReLU mask + reductions) and head gradients
int launch_dx_tc(const float* dpre, const float* u, const float* W_ih, const float* Wt, const float* bt,
                 const float* h, const float* dhead, float* partials, int nparts, long long P,
                 long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                 float* du, int B, int T, int Fin, int NH, void* stream);
// xproj_tc.cu: xproj[b,t,p] = bias + ReLU(torso) . W_ih^T for all rows, gate-interleaved columns
int launch_xproj_tc(const float* u, const float* Wt, const float* bt, const float* W_ih, const float* b_ih, const float* b_hh,
                    float* xproj, int B, int T, int Fin, int sms, void* stream);
}  // namespace b200rl
