// train.inl — empirical-Bayes training of the GP regression hyper-parameters (t, noise variance) on a spectrum handle
// (included by capi.cu inside its anonymous namespace).
//
//   reg_objective   negative_marginal_likelihood_regression_cpp / negative_log_posterior_regression_cpp, noise="same"
//                   (/root/reference/src/train.cpp:333-436; priors of PostOFDataReg, src/train.h:144-156)
//   mma_minimize    the optimiser behind nlopt_create(NLOPT_LD_MMA, 2) of train_regression_gp_cpp
//                   (src/train.cpp:557-671).  nloptr / NLopt is an un-vendored dependency of the reference (version
//                   unpinned, DESCRIPTION): restated from the published algorithm — Svanberg's CCSA with MMA
//                   approximations (SIAM J. Optim. 12, 2002) in NLopt's arrangement for bound constraints only.
//                   Parity with the reference is therefore to optimiser tolerance (xtol_rel = 1e-5), not bitwise.
//
// Everything n- and m-sized stays on the device: the objective needs only V1^T V1 (K x K), V1^T Y (K) and Y^T Y of the
// m training rows of the lifted eigenvectors (m > K, the Woodbury branch), or the m x K block itself (m <= K).
// What remains per evaluation is K x K (or m x m) dense algebra, K <= a few hundred: host C++ (the reference: Eigen LLT).

struct RegTrain {
  int m = 0, K = 0;
  double sigma = 1e-5;
  std::vector<double> ev;   // 1 - values[:K]
  std::vector<double> VtV;  // K x K (m > K)
  std::vector<double> b;    // V1^T Y (m > K)
  double yty = 0.0;
  std::vector<double> V;    // m x K row-major (m <= K)
  std::vector<double> Y;    // m (m <= K)
};

// statistics from the training rows V1 (row-major m_local x KK on the device, this rank's share from global row
// row_offset), eigenvalues `values` (as exported: the Laplacian spectrum is 1 - values)
RegTrain reg_train_from_rows(Ctx* c, const double* V1, int KK, int64_t m_local, int64_t row_offset, const double* Ydev,
                             int64_t m_total, int K, double sigma, const std::vector<double>& values) {
  need(K >= 1 && K <= KK, "K exceeds the number of computed eigenpairs");
  need(m_total >= 1 && m_total < INT32_MAX, "bad number of training rows");
  RegTrain T;
  T.m = (int)m_total;
  T.K = K;
  T.sigma = sigma;
  T.ev.resize(K);
  for (int k = 0; k < K; ++k) T.ev[k] = 1.0 - values[k];
  if (m_total > K) {
    const size_t words = (size_t)KK * KK + KK + 1;
    DevBuf<double> Gg(words);
    Gg.zero(c->stream);
    gram_small_run(c, V1, Ydev, m_local, KK, Gg.p, Gg.p + (size_t)KK * KK);
    if (m_local > 0) gemv_run(c, Ydev, Ydev, 1, (int)m_local, Gg.p + (size_t)KK * KK + KK);
    comm_allreduce_f64(c, Gg.p, words);
    std::vector<double> h(words);
    Gg.download(h.data(), words, c->stream);
    sync(c);
    T.VtV.resize((size_t)K * K);
    T.b.resize(K);
    for (int j = 0; j < K; ++j) {
      for (int i = 0; i < K; ++i) T.VtV[i + (size_t)K * j] = h[i + (size_t)KK * j];
      T.b[j] = h[(size_t)KK * KK + j];
    }
    T.yty = h[(size_t)KK * KK + KK];
  } else {
    const int m = (int)m_total;
    DevBuf<double> Vall((size_t)m * KK + m);
    Vall.zero(c->stream);
    if (m_local > 0) {
      FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)row_offset * KK, V1, sizeof(double) * m_local * KK,
                                cudaMemcpyDeviceToDevice, c->stream));
      FLGP_CUDA(cudaMemcpyAsync(Vall.p + (size_t)m * KK + row_offset, Ydev, sizeof(double) * m_local,
                                cudaMemcpyDeviceToDevice, c->stream));
    }
    comm_allreduce_f64(c, Vall.p, (size_t)m * KK + m);
    std::vector<double> h((size_t)m * KK + m);
    Vall.download(h.data(), h.size(), c->stream);
    sync(c);
    T.V.resize((size_t)m * K);
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < K; ++k) T.V[(size_t)i * K + k] = h[(size_t)i * KK + k];
    T.Y.assign(h.begin() + (size_t)m * KK, h.end());
  }
  return T;
}

RegTrain reg_train_prepare(flgp_spectrum* sp, const double* Ydev, int64_t m_total, int K, double sigma) {
  Ctx* c = sp->c;
  need(K >= 1 && K <= sp->K, "K exceeds the number of computed eigenpairs");
  need(m_total >= 1 && m_total <= sp->n_total && m_total < INT32_MAX, "bad number of training rows");
  const int r = sp->r, KK = sp->K;
  const int64_t m_local = std::max<int64_t>(0, std::min<int64_t>(sp->n_local, m_total - sp->row_offset));
  DevBuf<double> V1((size_t)std::max<int64_t>(m_local * KK, 1));
  lift_rows_run(c, r, sp->Zj.p, sp->Zx.p, sp->w.p, sp->Wm.p, KK, nullptr, m_local, V1.p, KK, false);
  return reg_train_from_rows(c, V1.p, KK, m_local, sp->row_offset, Ydev, m_total, K, sigma, sp->values);
}

// inverse from a lower Cholesky factor (column-major n x n): returns A^-1 (full, symmetric)
std::vector<double> chol_inverse(const std::vector<double>& L, int n) {
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) I[i + (size_t)n * i] = 1.0;
  chol_solve(L, n, I.data(), n);
  return I;
}

// objective value and gradient at x = (t, noise); posterior adds the priors.  Returns +inf when the system is not
// positive definite (the reference's LLT would silently produce NaN there).
double reg_objective(const RegTrain& T, const double* x, double* grad, bool posterior) {
  const int m = T.m, K = T.K;
  const double t = x[0], noise = x[1], ns = noise + T.sigma;
  double g0 = 0.0, g1 = 0.0, nmll = 0.0;
  if (m <= K) {
    // C = V diag(exp(-t ev)) V^T + (sigma + noise) I     (src/train.cpp:347-372)
    std::vector<double> C((size_t)m * m), lam(K), dlam(K);
    for (int k = 0; k < K; ++k) {
      lam[k] = std::exp(-t * T.ev[k]);
      dlam[k] = -T.ev[k] * lam[k];
    }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int k = 0; k < K; ++k) a += (T.V[(size_t)i * K + k] * lam[k]) * T.V[(size_t)j * K + k];
        C[i + (size_t)m * j] = a + (i == j ? (T.sigma + noise) : 0.0);
      }
    std::vector<double> L = C;
    if (!chol_lower(L, m)) return INFINITY;
    std::vector<double> alpha = T.Y;
    chol_solve(L, m, alpha.data(), 1);
    if (grad) {
      const std::vector<double> Ci = chol_inverse(L, m);
      // U = alpha alpha^T - C^-1;  grad_t = V diag(dlam) V^T
      double s0 = 0.0, tr = 0.0;
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
          double gt = 0.0;
          for (int k = 0; k < K; ++k) gt += (T.V[(size_t)j * K + k] * dlam[k]) * T.V[(size_t)i * K + k];  // grad_t(j, i)
          s0 += (alpha[i] * alpha[j] - Ci[i + (size_t)m * j]) * gt;
        }
      for (int i = 0; i < m; ++i) tr += alpha[i] * alpha[i] - Ci[i + (size_t)m * i];
      g0 = -0.5 * s0;
      g1 = -0.5 * tr;
    }
    double ya = 0.0, ld = 0.0;
    for (int i = 0; i < m; ++i) {
      ya += T.Y[i] * alpha[i];
      ld += std::log(L[i + (size_t)m * i] + 1e-9);
    }
    nmll = 0.5 * ya + ld;
  } else {
    // Woodbury branch (src/train.cpp:373-417) through the K-sized statistics:
    //   Q = ls VtV ls + ns I,  w = ls Q^-1 ls b,  V^T alpha = (b - VtV w) / ns,  Y^T alpha = (yty - b.w) / ns,
    //   alpha^T alpha = (yty - 2 b.w + w^T VtV w) / ns^2
    std::vector<double> ls(K), A(K), Q((size_t)K * K);
    for (int k = 0; k < K; ++k) {
      ls[k] = std::exp(-0.5 * t * T.ev[k]) + 0.0;
      A[k] = -T.ev[k] * (std::exp(-t * T.ev[k]) + 0.0) + 0.0;
    }
    for (int j = 0; j < K; ++j)
      for (int i = 0; i < K; ++i) Q[i + (size_t)K * j] = (ls[i] * T.VtV[i + (size_t)K * j]) * ls[j] + (i == j ? ns : 0.0);
    std::vector<double> L = Q;
    if (!chol_lower(L, K)) return INFINITY;
    std::vector<double> w(K), Gw(K);
    for (int k = 0; k < K; ++k) w[k] = ls[k] * T.b[k];
    chol_solve(L, K, w.data(), 1);
    for (int k = 0; k < K; ++k) w[k] *= ls[k];
    double bw = 0.0, wGw = 0.0;
    for (int i = 0; i < K; ++i) {
      double a = 0.0;
      for (int j = 0; j < K; ++j) a += T.VtV[i + (size_t)K * j] * w[j];
      Gw[i] = a;
      bw += T.b[i] * w[i];
      wGw += w[i] * a;
    }
    if (grad) {
      const std::vector<double> Qi = chol_inverse(L, K);
      double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
      for (int k = 0; k < K; ++k) {
        const double vta = (T.b[k] - Gw[k]) / ns;
        s1 += vta * A[k] * vta;
        s2 += A[k] * T.VtV[k + (size_t)K * k];
      }
      // s3 = sum_ij (Qi ls VtV)(i,j) (A VtV ls)(j,i);  s4 = sum_ij Qi(i,j) (ls VtV ls)(j,i)
      std::vector<double> P((size_t)K * K);  // P = Qi * (ls VtV)
      for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i) {
          double a = 0.0;
          for (int k = 0; k < K; ++k) a += Qi[k + (size_t)K * i] * (ls[k] * T.VtV[k + (size_t)K * j]);  // Qi symmetric
          P[i + (size_t)K * j] = a;
        }
      for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i) {
          s3 += P[i + (size_t)K * j] * ((A[j] * T.VtV[j + (size_t)K * i]) * ls[i]);
          s4 += Qi[i + (size_t)K * j] * ((ls[j] * T.VtV[j + (size_t)K * i]) * ls[i]);
        }
      g0 = -0.5 * s1 + 0.5 / ns * s2 - 0.5 / ns * s3;
      const double aa = (T.yty - 2.0 * bw + wGw) / (ns * ns);
      g1 = -0.5 * aa + 0.5 / ns * ((double)m - s4);
    }
    double ld = 0.0;
    for (int i = 0; i < K; ++i) ld += std::log(L[i + (size_t)K * i] + 1e-9);
    nmll = 0.5 * (T.yty - bw) / ns + ld + 0.5 * (double)(m - K) * std::log(ns);
  }
  if (grad) {
    if (std::fabs(g1) >= 10.0) g1 = g1 / std::fabs(g1) * 10.0;  // the reference's gradient clipping
    grad[0] = g0;
    grad[1] = g1;
  }
  if (posterior) {  // src/train.cpp:333-350, PostOFDataReg defaults
    const double p = 1.0, q = 10.0, tau = 2.0, al = 1e-1, be = 1e-3;
    nmll += p * std::log(t + 1e-9) + std::pow(t / tau, -q);
    nmll += (al + 1.0) * std::log(ns) + be / ns;
    if (grad) {
      grad[0] += p / (t + 1e-9) - (q / tau) * std::pow(t / tau, -q - 1.0);
      grad[1] += (al + 1.0) / ns - be / (ns * ns);
    }
  }
  return nmll;
}

// CCSA-MMA, bound constraints only.  f(x, grad) -> value.  Returns the number of evaluations.
template <class Fn>
int mma_minimize(int n, Fn&& f, const double* lb, const double* ub, double* x, double* minf_out, double xtol_rel,
                 int maxeval) {
  std::vector<double> sigma(n), dfdx(n), dcur(n), xcur(x, x + n), xprev(x, x + n), xprevprev(x, x + n);
  for (int j = 0; j < n; ++j) sigma[j] = (std::isinf(ub[j]) || std::isinf(lb[j])) ? 1.0 : 0.5 * (ub[j] - lb[j]);
  double rho = 1.0;
  double minf = f(x, dfdx.data());
  int nev = 1, k = 0;
  while (true) {
    ++k;
    if (k > 1) xprevprev = xprev;
    xprev = xcur;
    while (true) {
      double gval = minf, wval = 0.0;
      for (int j = 0; j < n; ++j) {
        xcur[j] = x[j];
        if (sigma[j] == 0.0) continue;
        const double s2 = sigma[j] * sigma[j];
        const double v = std::fabs(dfdx[j]) * sigma[j] + 0.5 * rho;
        const double u = dfdx[j] * s2;
        const double q = u / (v * sigma[j]);
        double dx = (u / v) / (-1.0 - std::sqrt(std::fabs(1.0 - q * q)));
        double xj = x[j] + dx;
        if (xj > ub[j]) xj = ub[j];
        else if (xj < lb[j]) xj = lb[j];
        if (xj > x[j] + 0.9 * sigma[j]) xj = x[j] + 0.9 * sigma[j];
        else if (xj < x[j] - 0.9 * sigma[j]) xj = x[j] - 0.9 * sigma[j];
        xcur[j] = xj;
        dx = xj - x[j];
        const double dx2 = dx * dx, den = 1.0 / (s2 - dx2);
        gval += (dfdx[j] * s2 * dx + v * dx2) * den;
        wval += 0.5 * dx2 * den;
      }
      const double fcur = f(xcur.data(), dcur.data());
      ++nev;
      const bool inner_done = gval >= fcur;
      if (fcur < minf) {
        minf = fcur;
        for (int j = 0; j < n; ++j) {
          x[j] = xcur[j];
          dfdx[j] = dcur[j];
        }
      }
      if (nev >= maxeval) {
        *minf_out = minf;
        return nev;
      }
      if (inner_done) break;
      if (fcur > gval) rho = std::min(10.0 * rho, 1.1 * (rho + (fcur - gval) / wval));
      else if (!(fcur == fcur)) rho = 10.0 * rho;  // NaN objective: shrink the step
    }
    double dn = 0.0, xn = 0.0;
    for (int j = 0; j < n; ++j) {
      dn += std::fabs(xcur[j] - xprev[j]);
      xn += std::fabs(xcur[j]);
    }
    if (dn <= xtol_rel * xn) break;
    rho = std::max(0.1 * rho, 1e-5);
    if (k > 1)
      for (int j = 0; j < n; ++j) {
        const double dx2 = (xcur[j] - xprev[j]) * (xprev[j] - xprevprev[j]);
        sigma[j] *= dx2 < 0 ? 0.7 : (dx2 > 0 ? 1.2 : 1.0);
        if (!std::isinf(ub[j]) && !std::isinf(lb[j])) {
          sigma[j] = std::min(sigma[j], 10.0 * (ub[j] - lb[j]));
          sigma[j] = std::max(sigma[j], 0.01 * (ub[j] - lb[j]));
        }
      }
  }
  *minf_out = minf;
  return nev;
}

// train_regression_gp_cpp, noise="same": x0 = (10, 1), lb = (1e-3, 1e-4), ub = +inf, xtol_rel = 1e-5.
// pars_io: in = start (or the defaults when NaN), out = optimum.  Returns obj = -minimum.
double train_regression(const RegTrain& T, bool posterior, double* pars_io, int* nevals) {
  double x[2] = {pars_io[0] == pars_io[0] ? pars_io[0] : 10.0, pars_io[1] == pars_io[1] ? pars_io[1] : 1.0};
  const double lb[2] = {1e-3, 1e-4}, ub[2] = {INFINITY, INFINITY};
  double minf = 0.0;
  const int nev = mma_minimize(2, [&](const double* xx, double* g) { return reg_objective(T, xx, g, posterior); }, lb, ub,
                               x, &minf, 1e-5, 1000);
  pars_io[0] = x[0];
  pars_io[1] = x[1];
  if (nevals) *nevals = nev;
  return -minf;
}

// ---- noise = "different": one noise variance per training row --------------------------------------------------------
//   reg_objective_diff  negative_marginal_likelihood_diff_noise_regression_cpp / negative_log_posterior_diff_noise_
//                       regression_cpp (/root/reference/src/train.cpp:438-556), x = (t, noise_1 .. noise_m), a single
//                       response column (q = 1).  m <= K: the m x m form; m > K: the Woodbury form through
//                       V^T Z^-1 V (K x K), Z = diag(noise_i + sigma).  The reference's clipping of the noise gradients to
//                       [-1, 1] in the Woodbury branch (:536-541) is reproduced; its value keeps the reference's terms
//                       (the source itself remarks that the value "is wrong").
//   train_regression_diff  train_regression_gp_cpp, noise = "different" (:588-611): MMA on m + 1 variables from
//                       x0 = (10, 1, .., 1), lb = (1e-3, 1e-4, ..), ub = +inf.  The reference reads m for these vectors
//                       through a pointer of the wrong struct type — undefined behaviour there;
//                       here m is the number of training rows, the evident intent.
//   reg_diff_coef       predict_regression_cpp, noisepar = "different" (src/Predict.cpp:76-113): both branches end in
//                       Y_pred = V_new (Lam (V1^T alpha)); returns coef = Lam V1^T alpha (K) for the folded mean.
struct RegTrainDiff {
  int m = 0, K = 0;
  double sigma = 1e-5;
  std::vector<double> ev;  // 1 - values[:K]
  std::vector<double> V;   // m x K row-major: the training rows of the eigenvectors
  std::vector<double> Y;   // m
};

// alpha of both branches (GPML algorithm 2.1 / its Woodbury form); optionally the factors the gradient needs.
// Returns false when a system is not positive definite.
bool reg_diff_alpha(const RegTrainDiff& T, const double* x, std::vector<double>& alpha, std::vector<double>* Lout,
                    std::vector<double>* Gout) {
  const int m = T.m, K = T.K;
  const double t = x[0];
  alpha.assign(m, 0.0);
  if (m <= K) {
    std::vector<double> lam(K), L((size_t)m * m);
    for (int k = 0; k < K; ++k) lam[k] = std::exp(-t * T.ev[k]);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int k = 0; k < K; ++k) a += (T.V[(size_t)i * K + k] * lam[k]) * T.V[(size_t)j * K + k];
        L[i + (size_t)m * j] = a + (i == j ? T.sigma + x[i + 1] : 0.0);
      }
    if (!chol_lower(L, m)) return false;
    alpha = T.Y;
    chol_solve(L, m, alpha.data(), 1);
    if (Lout) *Lout = std::move(L);
    return true;
  }
  std::vector<double> ls(K), zi(m), G((size_t)K * K, 0.0), u(K, 0.0), L((size_t)K * K);
  for (int k = 0; k < K; ++k) ls[k] = std::exp(-0.5 * t * T.ev[k]) + 0.0;
  for (int i = 0; i < m; ++i) zi[i] = 1.0 / (x[i + 1] + T.sigma);
  for (int i = 0; i < m; ++i) {  // G = V^T Z^-1 V,  u = V^T (Z^-1 Y)
    const double* v = &T.V[(size_t)i * K];
    const double zy = zi[i] * T.Y[i];
    for (int b = 0; b < K; ++b) {
      const double zb = zi[i] * v[b];
      u[b] += v[b] * zy;
      for (int a = 0; a < K; ++a) G[a + (size_t)K * b] += v[a] * zb;
    }
  }
  for (int b = 0; b < K; ++b)
    for (int a = 0; a < K; ++a) L[a + (size_t)K * b] = (ls[a] * G[a + (size_t)K * b]) * ls[b] + (a == b ? 1.0 : 0.0);
  if (!chol_lower(L, K)) return false;
  std::vector<double> w(K);
  for (int k = 0; k < K; ++k) w[k] = ls[k] * u[k];
  chol_solve(L, K, w.data(), 1);
  for (int i = 0; i < m; ++i) {
    const double* v = &T.V[(size_t)i * K];
    double a = 0.0;
    for (int k = 0; k < K; ++k) a += (v[k] * ls[k]) * w[k];
    alpha[i] = zi[i] * T.Y[i] - zi[i] * a;
  }
  if (Lout) *Lout = std::move(L);
  if (Gout) *Gout = std::move(G);
  return true;
}

double reg_objective_diff(const RegTrainDiff& T, const double* x, double* grad, bool posterior) {
  const int m = T.m, K = T.K;
  const double t = x[0];
  std::vector<double> alpha, L, G;
  if (!reg_diff_alpha(T, x, alpha, &L, &G)) {
    if (grad)
      for (int i = 0; i <= m; ++i) grad[i] = 0.0;
    return INFINITY;
  }
  double nmll = 0.0;
  std::vector<double> dlam(K);
  for (int k = 0; k < K; ++k) dlam[k] = -T.ev[k] * std::exp(-t * T.ev[k]);
  double ya = 0.0;
  for (int i = 0; i < m; ++i) ya += T.Y[i] * alpha[i];
  if (m <= K) {
    if (grad) {
      const std::vector<double> Ci = chol_inverse(L, m);
      double s0 = 0.0;
      for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
          double gt = 0.0;
          for (int k = 0; k < K; ++k) gt += (T.V[(size_t)j * K + k] * dlam[k]) * T.V[(size_t)i * K + k];  // grad_t(j, i)
          s0 += (alpha[i] * alpha[j] - Ci[i + (size_t)m * j]) * gt;
        }
      grad[0] = -0.5 * s0;
      for (int i = 0; i < m; ++i) grad[i + 1] = -0.5 * (alpha[i] * alpha[i] - Ci[i + (size_t)m * i]);
    }
    double ld = 0.0;
    for (int i = 0; i < m; ++i) ld += std::log(L[i + (size_t)m * i] + 1e-9);
    nmll = 0.5 * ya + ld;
  } else {
    std::vector<double> ls(K);
    for (int k = 0; k < K; ++k) ls[k] = std::exp(-0.5 * t * T.ev[k]) + 0.0;
    if (grad) {
      const std::vector<double> Qi = chol_inverse(L, K);
      std::vector<double> wa(K, 0.0);
      for (int i = 0; i < m; ++i)
        for (int k = 0; k < K; ++k) wa[k] += T.V[(size_t)i * K + k] * alpha[i];
      double s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int k = 0; k < K; ++k) {
        s1 += wa[k] * dlam[k] * wa[k];
        s2 += dlam[k] * G[k + (size_t)K * k];
      }
      std::vector<double> P((size_t)K * K);  // P = Qi (ls G)
      for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i) {
          double a = 0.0;
          for (int k = 0; k < K; ++k) a += Qi[k + (size_t)K * i] * (ls[k] * G[k + (size_t)K * j]);  // Qi symmetric
          P[i + (size_t)K * j] = a;
        }
      for (int j = 0; j < K; ++j)
        for (int i = 0; i < K; ++i) s3 += P[i + (size_t)K * j] * ((dlam[j] * G[j + (size_t)K * i]) * ls[i]);
      grad[0] = -0.5 * s1 + 0.5 * s2 - 0.5 * s3;
      std::vector<double> tmp(K), qt(K);
      for (int i = 0; i < m; ++i) {
        const double zi = 1.0 / (x[i + 1] + T.sigma);
        for (int k = 0; k < K; ++k) tmp[k] = (zi * T.V[(size_t)i * K + k]) * ls[k];
        double quad = 0.0;
        for (int b = 0; b < K; ++b) {
          double a = 0.0;
          for (int k = 0; k < K; ++k) a += tmp[k] * Qi[k + (size_t)K * b];
          quad += a * tmp[b];
        }
        double g = -0.5 * (alpha[i] * alpha[i]) + 0.5 * (zi - quad);
        if (std::fabs(g) >= 1.0) g = g / std::fabs(g) * 1.0;  // the reference's gradient clipping (:536-541)
        grad[i + 1] = g;
      }
    }
    double ld = 0.0, lz = 0.0;
    for (int i = 0; i < K; ++i) ld += std::log(L[i + (size_t)K * i] + 1e-9);
    for (int i = 0; i < m; ++i) lz += std::log((x[i + 1] + T.sigma) + 1e-9);
    nmll = 0.5 * ya + ld + 0.5 * lz;
  }
  if (posterior) {  // src/train.cpp:438-458, PostOFDataReg defaults
    const double p = 1.0, q = 10.0, tau = 2.0, al = 1e-1, be = 1e-3;
    nmll += p * std::log(t + 1e-9) + std::pow(t / tau, -q);
    if (grad) grad[0] += p / (t + 1e-9) - (q / tau) * std::pow(t / tau, -q - 1.0);
    double pr1 = 0.0;
    for (int i = 0; i < m; ++i) {
      const double ns = x[i + 1] + T.sigma;
      pr1 += ((al + 1.0) * std::log(ns) + be / ns) / m;
      if (grad) grad[i + 1] += ((al + 1.0) / ns - be / (ns * ns)) / m;
    }
    nmll += pr1;
  }
  return nmll;
}

// x_io: m + 1 values, NaN entries take the reference's start (10, 1, .., 1).  Returns obj = -minimum.
double train_regression_diff(const RegTrainDiff& T, bool posterior, double* x_io, int* nevals) {
  const int n = T.m + 1;
  std::vector<double> lb(n, 1e-4), ub(n, INFINITY);
  lb[0] = 1e-3;
  for (int i = 0; i < n; ++i)
    if (!(x_io[i] == x_io[i])) x_io[i] = i == 0 ? 10.0 : 1.0;
  double minf = 0.0;
  const int nev = mma_minimize(n, [&](const double* xx, double* g) { return reg_objective_diff(T, xx, g, posterior); },
                               lb.data(), ub.data(), x_io, &minf, 1e-5, 1000);
  if (nevals) *nevals = nev;
  return -minf;
}

// coef = Lam V1^T alpha  (K): Y_pred = V_new coef in both branches of predict_regression_cpp, noisepar = "different"
bool reg_diff_coef(const RegTrainDiff& T, const double* x, std::vector<double>& coef) {
  std::vector<double> alpha;
  if (!reg_diff_alpha(T, x, alpha, nullptr, nullptr)) return false;
  coef.assign(T.K, 0.0);
  for (int i = 0; i < T.m; ++i)
    for (int k = 0; k < T.K; ++k) coef[k] += T.V[(size_t)i * T.K + k] * alpha[i];
  for (int k = 0; k < T.K; ++k) coef[k] *= std::exp(-x[0] * T.ev[k]);
  return true;
}

// ---- binary GP classifier: training of the diffusion time t (second half) -----------------------
//   laplace_mll       marginal_log_likelihood_logit_la_cpp (/root/reference/src/train.cpp:716-760): Newton iterations for
//                     the posterior mode (GPML algorithm 3.1) from f = 0, stop when |f - f_new|_1 < tol; the value uses
//                     `a` and chol(B) of the LAST Newton step, as the reference does.
//   logit_objective   negative_marginal_likelihood_logit_cpp / negative_log_posterior_logit_cpp (src/train.cpp:14-36):
//                     C = V1 diag(exp(-t (1 - values))) V1^T + sigma I on the m labelled rows; prior of PostOFData
//                     (src/train.h:129-141): p log(t + 1e-9) + (t / tau)^(-q), p = 1e-2, q = 10, tau = 2.
//   cobyla_minimize_1d  the optimiser behind nlopt_create(NLOPT_LN_COBYLA, 1) of train_lae_logit_gp_cpp
//                     (src/train.cpp:38-71): t0 = 10, lb = 1e-3, ub = +inf, xtol_rel = 1e-4.  NLopt is an un-vendored
//                     dependency (parity unpinned); this is Powell's COBYLA iteration written out for ONE variable: a
//                     two-point simplex, the linear model through it, a trust-region step of length rho from the best
//                     vertex (clipped to the bounds), rho kept while the step pays off and divided by 10 otherwise,
//                     from rhobeg (NLopt's default initial step: 0.75 (t0 - lb) here) down to rhoend = xtol_rel rhobeg.
//                     Trained t agrees with NLopt / scipy COBYLA to the optimiser's tolerance on a unimodal objective.
struct LogitTrain {
  int m = 0, K = 0;
  double sigma = 1e-3;
  bool posterior = true;
  double p = 1e-2, q = 10.0, tau = 2.0;
  std::vector<double> V;   // m x K row-major: the labelled rows of the lifted eigenvectors
  std::vector<double> ev;  // 1 - values[:K]
  std::vector<double> Y, N;
};

double laplace_mll(const std::vector<double>& C, const double* Y, const double* N, int m, double tol, int max_iter) {
  std::vector<double> f(m, 0.0), pi(m), W(m), sw(m), b(m), cb(m), a(m, 0.0), fn(m), B((size_t)m * m);
  double logdet = 0.0;
  for (int iter = 0; iter < max_iter; ++iter) {
    for (int i = 0; i < m; ++i) {
      pi[i] = 1.0 / (1.0 + std::exp(-f[i]));
      W[i] = N[i] * pi[i] * (1.0 - pi[i]);
      sw[i] = std::sqrt(W[i]);
    }
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) B[i + (size_t)m * j] = (sw[i] * C[i + (size_t)m * j]) * sw[j] + (i == j ? 1.0 : 0.0);
    if (!chol_lower(B, m)) fail(2, "classification: the Newton system is not positive definite");
    logdet = 0.0;
    for (int i = 0; i < m; ++i) logdet += std::log(B[i + (size_t)m * i] + 1e-9);
    for (int i = 0; i < m; ++i) b[i] = W[i] * f[i] + Y[i] * (1.0 - pi[i]) + (N[i] - Y[i]) * (-pi[i]);
    // C b and C a as column sweeps (contiguous; every row still adds its terms in ascending j: the same bits)
    std::fill(fn.begin(), fn.end(), 0.0);
    for (int j = 0; j < m; ++j) {
      const double bj = b[j];
      const double* cj = &C[(size_t)m * j];
      for (int i = 0; i < m; ++i) fn[i] += cj[i] * bj;
    }
    for (int i = 0; i < m; ++i) cb[i] = sw[i] * fn[i];
    chol_solve(B, m, cb.data(), 1);
    for (int i = 0; i < m; ++i) a[i] = b[i] - sw[i] * cb[i];
    double diff = 0.0;
    std::fill(fn.begin(), fn.end(), 0.0);
    for (int j = 0; j < m; ++j) {
      const double aj = a[j];
      const double* cj = &C[(size_t)m * j];
      for (int i = 0; i < m; ++i) fn[i] += cj[i] * aj;
    }
    for (int i = 0; i < m; ++i) diff += std::fabs(f[i] - fn[i]);
    f = fn;
    if (diff < tol) break;
  }
  double amll = 0.0;
  for (int i = 0; i < m; ++i) amll += a[i] * f[i];
  amll *= -0.5;
  for (int i = 0; i < m; ++i) {
    const double p1 = 1.0 / (1.0 + std::exp(-f[i]));
    amll += Y[i] * std::log(p1) + (N[i] - Y[i]) * std::log(1.0 - p1);
  }
  return amll - logdet;
}

double logit_objective(const LogitTrain& T, double t) {
  const int m = T.m, K = T.K;
  std::vector<double> lam(K), C((size_t)m * m);
  for (int k = 0; k < K; ++k) lam[k] = std::exp(-t * T.ev[k]);
  std::vector<double> VL((size_t)m * K);  // rows scaled by lam once: the product (v_ik lam_k) v_jk keeps its two roundings
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < K; ++k) VL[(size_t)i * K + k] = T.V[(size_t)i * K + k] * lam[k];
  const int Tn = m >= 256 ? std::min(host_threads(), m / 64) : 1;
  host_parallel(Tn, [&](int q, int step) {  // column j by one thread: the same sums for any thread count
    for (int j = q; j < m; j += step)
      for (int i = j; i < m; ++i) {
        double acc = 0.0;
        const double* vi = &VL[(size_t)i * K];
        const double* vj = &T.V[(size_t)j * K];
        for (int k = 0; k < K; ++k) acc += vi[k] * vj[k];
        C[i + (size_t)m * j] = C[j + (size_t)m * i] = acc + (i == j ? T.sigma : 0.0);
      }
  });
  const double mll = laplace_mll(C, T.Y.data(), T.N.data(), m, 1e-5, 100);
  double pr = 0.0;
  if (T.posterior) pr = T.p * std::log(t + 1e-9) + std::pow(t / T.tau, -T.q);
  return -mll + pr;
}

template <class F>
double cobyla_minimize_1d(F&& f, double x0, double lb, double ub, double xtol_rel, int maxeval, double* fmin_out,
                          int* nevals_out) {
  // NLopt's default initial step for one bounded variable (nlopt_set_default_initial_step)
  double step = HUGE_VAL;
  if (std::isfinite(ub) && std::isfinite(lb) && (ub - lb) * 0.25 < step && ub > lb) step = (ub - lb) * 0.25;
  if (std::isfinite(ub) && ub - x0 < step && ub > x0) step = (ub - x0) * 0.75;
  if (std::isfinite(lb) && x0 - lb < step && x0 > lb) step = (x0 - lb) * 0.75;
  if (!std::isfinite(step)) step = std::fabs(x0);
  if (!(step > 0.0)) step = 1.0;
  const double rhobeg = step, rhoend = xtol_rel * rhobeg;
  auto clip = [&](double x) { return std::min(ub, std::max(lb, x)); };
  int nev = 0;
  double xa = clip(x0), fa = f(xa);
  ++nev;
  double xb = (xa + rhobeg <= ub) ? xa + rhobeg : xa - rhobeg, fb = f(clip(xb));
  xb = clip(xb);
  ++nev;
  if (fb < fa) {
    std::swap(xa, xb);
    std::swap(fa, fb);
  }
  double rho = rhobeg;
  while (nev < maxeval) {
    double dir;  // descent direction of the linear model through the simplex
    if (xb != xa && fb != fa) dir = ((fb - fa) / (xb - xa) > 0.0) ? -1.0 : 1.0;
    else dir = (xb > xa) ? -1.0 : 1.0;
    double xt = clip(xa + dir * rho);
    if (xt == xa) xt = clip(xa - dir * rho);  // pinned at a bound: the only move is inwards
    bool improved = false;
    if (xt != xa) {
      const double ft = f(xt);
      ++nev;
      if (ft < fa) {
        const double predicted = (xb != xa) ? std::fabs((fb - fa) / (xb - xa)) * std::fabs(xt - xa) : 0.0;
        improved = (fa - ft) >= 0.1 * predicted;  // the model pays off at this radius: keep rho
        xb = xa;
        fb = fa;
        xa = xt;
        fa = ft;
      } else {
        xb = xt;  // the far vertex moves to the trial point: the simplex stays within rho of the best vertex
        fb = ft;
      }
    }
    if (!improved) {
      if (rho <= rhoend) break;
      rho *= 0.1;
      if (rho <= 1.5 * rhoend) rho = rhoend;
    }
  }
  if (fmin_out) *fmin_out = fa;
  if (nevals_out) *nevals_out = nev;
  return xa;
}
