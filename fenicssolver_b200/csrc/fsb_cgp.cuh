// Host interface of the persistent CG kernel (fsb_cgp.cu), used by the Krylov driver in fsb_solve.cu.
#pragma once
#include "fsb_device.cuh"

struct CgpVectors {
  const double* dinv;
  double *u, *w, *p, *s, *x, *r;
};

// does the staged SpMV configuration of S have a persistent-CG instantiation?
bool fsb_cgp_supported(fsb_mat* S);
// Runs the iteration loop to convergence / maxit / breakdown in one cooperative launch on ctx->stream.  On entry r, u = M^-1 r,
// w = A u hold the start state and the mailboxes of pc.buf[pc.rank] hold (r.u, u.u) in MAIL_RZ and w.u in MAIL_PQ with sequence
// number seq_base; d_scalars[bb_slot] = |M^-1 b|^2.  On exit d_state = {1, iterations, outcome} and d_scalars[final_slot] =
// |M^-1 r|^2.  phase_ms (profile mode): time worker 0 spent in the update phase, the grid barrier, the SpMV phase and the
// wait for the reduced scalars.
int fsb_cgp_run(fsb_mat* A, fsb_mat* S, const CgpVectors& v, double rtol, double atol, int maxit, int bb_slot, int final_slot,
                const PeerComm& pc, unsigned long long seq_base, double* phase_ms);
