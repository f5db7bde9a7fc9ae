// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// Philox4x32-10 (Salmon et al., SC'11; published algorithm) and the noise
// layout the TVLQR replay draws from.  The reference draws from Julia's
// MersenneTwister (src/simulator.jl:5,10,22), which is not reproducible outside
// Julia; this repo freezes a counter-based stream instead so that any trial is
// reproducible from (seed, trial, step, stage).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
  uint32_t k[2] = {key_in[0], key_in[1]};
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

inline double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

// 9 scaled perturbations for one dynamics call (quirk Q7):
//   [0:3) omega noise  = randn(3)*(.38*pi/180)^2     (simulator.jl:5)
//   [3:6) q noise      = randn(3)*(1*pi/180)^2       (simulator.jl:10)
//   [6:9) B noise      = rand(3)*(1e-5)^2            (simulator.jl:22)
inline void tvlqr_noise(uint64_t seed, uint32_t trial, uint32_t step, uint32_t stage, double out[9]) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r0[4], r1[4], r2[4];
  const uint32_t c0[4] = {trial, step, stage, 0}, c1[4] = {trial, step, stage, 1}, c2[4] = {trial, step, stage, 2};
  philox4x32_10(c0, key, r0);
  philox4x32_10(c1, key, r1);
  philox4x32_10(c2, key, r2);
  auto bm = [](uint32_t a, uint32_t b, double& z0, double& z1) {
    const double rad = std::sqrt(-2.0 * std::log(u01(a)));
    const double ang = 2.0 * M_PI * u01(b);
    z0 = rad * std::cos(ang);
    z1 = rad * std::sin(ang);
  };
  double nrm[6];
  bm(r0[0], r0[1], nrm[0], nrm[1]);
  bm(r0[2], r0[3], nrm[2], nrm[3]);
  bm(r1[0], r1[1], nrm[4], nrm[5]);
  const double s_w = (.38 * M_PI / 180) * (.38 * M_PI / 180);
  const double s_q = (1 * M_PI / 180) * (1 * M_PI / 180);
  const double s_b = (1E-5) * (1E-5);
  for (int i = 0; i < 3; ++i) out[i] = nrm[i] * s_w;
  for (int i = 0; i < 3; ++i) out[3 + i] = nrm[3 + i] * s_q;
  out[6] = u01(r1[2]) * s_b;
  out[7] = u01(r1[3]) * s_b;
  out[8] = u01(r2[0]) * s_b;
}

}  // namespace orc
