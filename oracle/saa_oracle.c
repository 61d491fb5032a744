/*
 * saa_oracle.c -- plain-C CPU restatement of the drone SAA linearize+assemble
 * step.  TEST INFRASTRUCTURE / CPU BASELINE ONLY: never linked into or called
 * by the product (riskaversetrajopt_b200); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it.  Parity is unpinned by the
 * reference (it ships no golden vectors); this file is pinned against
 * oracle_a/oracle_b by tests/test_oracle_c.py.
 *
 * Follows, per sample (reference drone/drone_risk.py):
 *   rollout                 :139-155  (Euler-Maruyama, double sqrt(dt) quirk)
 *   b, sigma                :122-137
 *   obstacle constraints    :169-213  g[o,k] = 1 - (p_k-c_o)^T Q_o (p_k-c_o), k = 1..S
 *   Jacobian wrt controls   :239-280  (jacfwd there; closed-form 2x2 chains here)
 *   linearisation offsets   :271, :278
 *   sample-mean final rows  :294-296
 *   packing of sample rows  :353-364  (0.01 multiplier)
 * and writes the values directly at their CSC positions (what sp.csr_matrix +
 * sp.vstack(format='csc') at :419-420 would produce), so it scales to 10^6+
 * samples where the reference's dense O(M^2) matrix cannot exist.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC saa_oracle.c -o _build/libsaa_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S_MAX 64

typedef struct {
  double dt, beta, drag, kp, kv;
  double x_init[6], x_final[6];
  double obs_pos[3][2];
  double mult;      /* 0.01 (drone_risk.py:352) */
  double pad;       /* subtracted from the upper bounds (baseline :324-325) */
  double escale;    /* extra scale on Jacobian entries (relaxation :413-417) */
} oracle_drone_params;

int saa_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/*
 * col_off[a*(S-1)+j] : position in Ax of sample 0's sub-run of u column (j, axis a), a in {0,1}
 * ub                 : base pointer of the sample rows' upper bounds (60 per sample)
 * mean_sums          : 3*(S-1) + 3*S + 6 doubles: sums over samples of
 *                      d p_S^a/du_j (a<3,j<S-1) | d v_S^a/du_j (a<3,j<S) | linearisation offsets (6)
 * Z                  : per-sample max_{o,k} g (may be NULL)
 */
void saa_oracle_drone_assemble(int64_t M, int S, const double *masses, const double *DWs,
                               const double *obs_Qs, const double *us,
                               const oracle_drone_params *P, double *Ax, const int64_t *col_off,
                               double *ub, double *mean_sums, double *Z) {
  const int NRED = 3 * (S - 1) + 3 * S + 6;
  const int FIN_V = 3 * (S - 1), VAL = FIN_V + 3 * S;
  const double dt = P->dt, sqdt = sqrt(P->dt);
  int nthreads = saa_oracle_threads();
  double *partial = (double *)calloc((size_t)nthreads * NRED, sizeof(double));
#pragma omp parallel
  {
#ifdef _OPENMP
    double *acc = partial + (size_t)omp_get_thread_num() * NRED;
#else
    double *acc = partial;
#endif
    double p[3][S_MAX + 1], v[3][S_MAX + 1], a22[3][S_MAX], tp[3][S_MAX + 1], tv[3];
#pragma omp for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
      const double m = masses[i];
      const double *dW = DWs + i * S * 6;
      const double *Q = obs_Qs + i * 27;
      const double a21 = -P->kp * dt / m;
      for (int a = 0; a < 3; ++a) {
        p[a][0] = P->x_init[a]; v[a][0] = P->x_init[3 + a]; tp[a][0] = 0.0; tv[a] = 0.0;
        for (int k = 0; k < S; ++k) {
          const double pk = p[a][k], vk = v[a][k], u = us[k * 3 + a];
          const double acc_ = (u - P->kp * pk - P->kv * vk) / m - P->drag * fabs(vk) * vk / m;
          a22[a][k] = 1.0 - dt * (P->kv + 2.0 * P->drag * fabs(vk)) / m;
          const double ntp = tp[a][k] + dt * tv[a];
          tv[a] = a21 * tp[a][k] + a22[a][k] * tv[a] + dt / m * u;   /* tangent along u itself */
          tp[a][k + 1] = ntp;
          p[a][k + 1] = pk + dt * vk;
          v[a][k + 1] = vk + dt * acc_ + sqdt * (P->beta / m) * dW[k * 6 + 3 + a];
        }
        acc[VAL + a] += -(p[a][S] - P->x_final[a]) + tp[a][S];
        acc[VAL + 3 + a] += -(v[a][S] - P->x_final[3 + a]) + tv[a];
      }
      /* upper bounds of the sample rows and Z_i */
      double zmax = -INFINITY;
      for (int o = 0; o < 3; ++o) {
        const double qx = Q[o * 9 + 0], qy = Q[o * 9 + 4];
        for (int k = 1; k <= S; ++k) {
          const double dx = p[0][k] - P->obs_pos[o][0], dy = p[1][k] - P->obs_pos[o][1];
          const double g = 1.0 - (qx * dx * dx + qy * dy * dy);
          const double gdu_u = -2.0 * qx * dx * tp[0][k] - 2.0 * qy * dy * tp[1][k];
          ub[i * 3 * S + o * S + (k - 1)] = P->mult * (-g + gdu_u) - P->pad;
          if (g > zmax) zmax = g;
        }
      }
      if (Z) Z[i] = zmax;
      /* sensitivity chains */
      for (int a = 0; a < 3; ++a) {
        for (int j = 0; j < S; ++j) {
          double sp = 0.0, sv = dt / m;
          const int L = S - 1 - j;
          double *dst = (a < 2 && L > 0) ? Ax + col_off[a * (S - 1) + j] + i * 3 * L : NULL;
          for (int k = j + 1; k < S; ++k) {
            const double nsp = sp + dt * sv;
            sv = a21 * sp + a22[a][k] * sv;
            sp = nsp;
            if (dst) {
              const int kk = k - j - 1;
              for (int o = 0; o < 3; ++o) {
                const double q = Q[o * 9 + a * 4];
                dst[o * L + kk] = P->mult * P->escale * (-2.0 * q * (p[a][k + 1] - P->obs_pos[o][a])) * sp;
              }
            }
          }
          if (j < S - 1) acc[a * (S - 1) + j] += sp;
          acc[FIN_V + a * S + j] += sv;
        }
      }
    }
  }
  for (int r = 0; r < NRED; ++r) {
    double s = 0.0;
    for (int t = 0; t < nthreads; ++t) s += partial[(size_t)t * NRED + r];
    mean_sums[r] = s;
  }
  free(partial);
}
