// HOST LANE-EMULATOR -- TEST INFRASTRUCTURE ONLY.
// Compiles the *same* team-cooperative AL-iLQR source the CUDA kernel K3 uses
// (tortoisesat.jl_b200/csrc/ilqr_solver.cuh, ilqr_math.cuh) for the CPU, with the
// warp intrinsics emulated by 8 OS threads + barriers, so the kernel's logic,
// indexing and synchronisation points can be checked against the oracle here
// (no GPU in the build container).  Never linked into the product library.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../tortoisesat.jl_b200/csrc/ilqr_solver.cuh"
#include "../../tortoisesat.jl_b200/csrc/tvlqr_solver.cuh"

using namespace ts;

// Sense-reversing barrier that spins briefly and then yields: the emulated teams synchronise three times per
// knot, and a futex sleep/wake per synchronisation (std::barrier) made the CPU suite system-time bound.
struct SpinBarrier {
  explicit SpinBarrier(int n_) : n(n_) {}
  void arrive_and_wait() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 200) std::this_thread::yield();
    }
  }
  std::atomic<int> count{0}, gen{0};
  const int n;
};

template <int W_>
struct TeamSharedT {
  SpinBarrier bar{W_};
  double xch[W_];
  unsigned bits[W_];
  std::vector<double> sm = std::vector<double>(SmL<W_>::TOTAL, 0.0);
};
template <int W_>
struct CpuTeamT {
  static constexpr int W = W_;
  TeamSharedT<W_>* sh;
  int ln;
  int lane() const { return ln; }
  double* smem() const { return sh->sm.data(); }
  void sync() const { sh->bar.arrive_and_wait(); }
  double bcast(double v, int src) const {
    sh->xch[ln] = v;
    sync();
    const double r = sh->xch[src];
    sync();
    return r;
  }
  unsigned ballot(bool p) const {
    sh->bits[ln] = p ? 1u : 0u;
    sync();
    unsigned r = 0;
    for (int i = 0; i < W_; ++i) r |= sh->bits[i] << i;
    sync();
    return r;
  }
  void stage16(double* dst, const double* src, int n16) const { memcpy(dst, src, (size_t)n16 * 16); }
  void prefetch_l2(const void*) const {}
  void stage_commit() const {}
  void stage_wait(int) const {}
  double sum(double v) const {
    for (int off = W_ / 2; off >= 1; off >>= 1) {
      sh->xch[ln] = v;
      sync();
      v += sh->xch[ln ^ off];
      sync();
    }
    return v;
  }
  double max(double v) const {
    for (int off = W_ / 2; off >= 1; off >>= 1) {
      sh->xch[ln] = v;
      sync();
      v = fmax(v, sh->xch[ln ^ off]);
      sync();
    }
    return v;
  }
};

template <int W_>
struct CpuQuatTeamT : CpuTeamT<W_> {   // the quaternion-aware variant (ts_ilqr_opts.quat_error)
  static constexpr bool QUAT = true;
};

template <int W_, class TeamT = CpuTeamT<W_>>
static void solve_with_width(const TrialIn& in, const ts_ilqr_opts_dev* opts, int64_t N, double* X, double* U, double* K,
                             ts_trial_outcome_dev* out) {
  const size_t per_slot = (size_t)9 * N * 10;
  std::vector<double> xu(4 * per_slot), kd((size_t)N * 24), lam((size_t)N * 6), clk((size_t)N), bk((size_t)N * 10);
  TrialWork w;
  w.xu = xu.data();
  w.xu_warp = xu.data();
  w.slot_stride = (long long)per_slot;
  w.kd = kd.data();
  w.lam = lam.data();
  w.clk = clk.data();
  w.bk = bk.data();
  w.Nmax = N;
  TeamSharedT<W_> sh;
  ts_trial_outcome_dev oc[W_];
  int cur[W_];
  std::vector<std::thread> th;
  for (int l = 0; l < W_; ++l)
    th.emplace_back([&, l]() {
      TeamT tm;
      tm.sh = &sh;
      tm.ln = l;
      alilqr_solve_team(tm, in, *opts, w, oc[l], cur[l]);
    });
  for (auto& t : th) t.join();
  *out = oc[0];
  const double* x = xu_buf<W_>(w, cur[0]);
  for (int64_t k = 0; k < N; ++k) {
    for (int i = 0; i < 7; ++i) X[k * 8 + i] = x[k * 10 + i];
    X[k * 8 + 7] = clk[k];
    if (k < N - 1) {
      for (int i = 0; i < 3; ++i) U[k * 3 + i] = x[k * 10 + 7 + i];
      if (K)
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 7; ++j) K[k * 24 + i * 8 + j] = kd[k * 24 + j * 3 + i];
          K[k * 24 + i * 8 + 7] = 0.0;
        }
    }
  }
}


extern "C" {

void hs_dyn_f(const double* J9, const double* x7, const double* u3, const double* Bn, double* dx7) {
  Inertia I;
  memcpy(I.J, J9, 72);
  inv3_gj(I.J, I.Jinv);
  dyn_f<0>(I, x7, u3, Bn, dx7);
}
void hs_rk3_jac7(const double* J9, const double* x7, const double* u3, const double* B1, const double* B2, const double* B3,
                 double dt, double* xn7, double* AB70) {
  Inertia I;
  memcpy(I.J, J9, 72);
  inv3_gj(I.J, I.Jinv);
  rk3_jac7<0>(I, x7, u3, B1, B2, B3, dt, xn7, AB70);
}
void hs_rk3_jac7_jvp(const double* J9, const double* x7, const double* u3, const double* B1, const double* B2, const double* B3,
                     double dt, double* colmajor70) {
  Inertia I;
  memcpy(I.J, J9, 72);
  inv3_gj(I.J, I.Jinv);
  rk3_jac7_jvp(I, x7, u3, B1, B2, B3, dt, colmajor70);
}
// the diagonal-inertia instantiation (mul_J<true> / mul_Jinv<true>: K3's *_diag_kernel), and its rollout step
void hs_rk3_jac7_jvp_diag(const double* J9, const double* x7, const double* u3, const double* B1, const double* B2, const double* B3,
                          double dt, double* colmajor70, double* xn7) {
  Inertia I;
  memcpy(I.J, J9, 72);
  inv3_gj(I.J, I.Jinv);
  rk3_jac7_jvp<true>(I, x7, u3, B1, B2, B3, dt, colmajor70);
  rk3_step7<0, true>(I, x7, u3, B1, B2, B3, dt, xn7);
}
void hs_rk4_jac7(const double* J9, const double* x7, const double* u3, const double* B1, const double* B2, const double* B3,
                 const double* B4, double dt, double* xn7, double* AB70) {
  Inertia I;
  memcpy(I.J, J9, 72);
  inv3_gj(I.J, I.Jinv);
  rk4_jac7<1>(I, x7, u3, B1, B2, B3, B4, dt, xn7, AB70);
}

// One trial, same argument meaning as the C ABI's ts_alilqr_solve_batch for n_trials = 1.
// width = 8: a narrow team (one of the four of a warp); width = 32: the wide (whole-warp) team.
void hs_alilqr_solve_w(int width, int64_t N, const double* x0, const double* xf, const double* Jmat, const double* Qd,
                       const double* Qfd, const double* Rd, const double* B_eci, int64_t B_rows, double index_scale,
                       double clock_rate, double dt, const double* U0, const ts_ilqr_opts_dev* opts, double* X, double* U, double* K,
                       ts_trial_outcome_dev* out) {
  TrialIn in;
  in.N = (int)N;
  in.dt = dt;
  for (int i = 0; i < 7; ++i) in.x0[i] = x0[i];
  in.clk0 = x0[7];
  for (int i = 0; i < 8; ++i) {
    in.xf[i] = xf[i];
    in.Qd[i] = Qd[i];
    in.Qfd[i] = Qfd[i];
  }
  for (int i = 0; i < 3; ++i) in.Rd[i] = Rd[i];
  memcpy(in.I.J, Jmat, 72);
  inv3_gj(in.I.J, in.I.Jinv);
  in.Bt = B_eci;
  in.B_rows = B_rows;
  in.index_scale = index_scale;
  in.clock_rate = clock_rate;
  in.U0 = U0;
  if (opts->quat_error && width == 32)
    solve_with_width<32, CpuQuatTeamT<32>>(in, opts, N, X, U, K, out);
  else if (opts->quat_error)
    solve_with_width<8, CpuQuatTeamT<8>>(in, opts, N, X, U, K, out);
  else if (width == 32)
    solve_with_width<32>(in, opts, N, X, U, K, out);
  else
    solve_with_width<8>(in, opts, N, X, U, K, out);
}
void hs_alilqr_solve(int64_t N, const double* x0, const double* xf, const double* Jmat, const double* Qd, const double* Qfd,
                     const double* Rd, const double* B_eci, int64_t B_rows, double index_scale, double clock_rate, double dt,
                     const double* U0, const ts_ilqr_opts_dev* opts, double* X, double* U, double* K, ts_trial_outcome_dev* out) {
  hs_alilqr_solve_w(8, N, x0, xf, Jmat, Qd, Qfd, Rd, B_eci, B_rows, index_scale, clock_rate, dt, U0, opts, X, U, K, out);
}

void hs_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
void hs_tvlqr_noise(uint64_t seed, uint32_t trial, uint32_t step, uint32_t stage, double* out9) {
  tvlqr_noise(seed, trial, step, stage, out9);
}
// One trial of the K4 source: gains + replay + slew-time rule.
int64_t hs_tvlqr(int64_t N, const double* X_lqr, const double* U_lqr, const double* x0, const double* Jmat, const double* B_eci,
                 int64_t B_rows, double index_scale, double clock_rate, double t_final, const double* q_final, uint32_t trial,
                 const ts_tvlqr_opts_dev* opts, const double* noise, double* X_sim, double* U_sim, double* dX, double* K,
                 double* slew_time) {
  TvlqrIn in;
  in.N = (int)N;
  in.X_lqr = X_lqr;
  in.U_lqr = U_lqr;
  for (int i = 0; i < 8; ++i) in.x0[i] = x0[i];
  memcpy(in.I.J, Jmat, 72);
  inv3_gj(in.I.J, in.I.Jinv);
  in.Bt = B_eci;
  in.B_rows = B_rows;
  in.index_scale = index_scale;
  in.clock_rate = clock_rate;
  in.noise = noise;
  in.trial = trial;
  for (int i = 0; i < 4; ++i) in.q_final[i] = q_final[i];
  in.t_final = t_final;
  in.time_step = opts->dt;
  in.trial_index_1based = (long long)trial + 1;
  ts_tvlqr_opts_dev o = *opts;
  o.tf = t_final;
  tvlqr_gains(in, o, K);
  return tvlqr_replay(in, o, K, X_sim, U_sim, dX, slew_time);
}

}  // extern "C"
