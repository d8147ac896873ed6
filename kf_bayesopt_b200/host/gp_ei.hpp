// gp_ei.hpp -- a small Gaussian-process / expected-improvement optimiser: the stand-in for the
// external `bayesopt` library the reference links (/usr/local/lib/libbayesopt.a, CMakeLists.txt:158;
// used at robot1d_bayesopt.cpp:28-31,225-264), which is not installed here.  It follows the reference's
// configuration (config/bayesParams.yaml): Gaussian-process surrogate, Matern-3/2 ARD kernel
// (kMaternARD3), expected-improvement criterion (cEI), n_init_samples Latin-hypercube points, then
// n_iterations sequential evaluations, kernel length-scales re-learned every n_iter_relearn
// iterations, the criterion maximised with n_inner_evaluation points per dimension.
// It is NOT a restatement of bayesopt's code: optimiser parity is unpinned; what is
// kept bit-for-bit is the evaluateSample contract -- one "noise" message out, one "cost" message back.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "kfbo_trial.hpp"

namespace kfbo {
namespace bo {

// The members of ReadParaBayes that the optimiser uses (include/read_para_bayes.h, config/bayesParams.yaml).
struct ReadParaBayes {
    int n_init_samples_ = 20;
    int n_iterations_ = 100;
    int random_seed_ = -1;
    int verbose_level_ = 1;
    double noise_ = 1e-8;
    int n_iter_relearn_ = 10;
    int n_inner_evaluation_ = 500;
    std::vector<double> lower_bound_{0.1, 0.01};
    std::vector<double> upper_bound_{5.0, 1.0};

    static ReadParaBayes FromYaml(const std::string &path)
    {
        const FlatYaml y = FlatYaml::Load(path);
        ReadParaBayes rb;
        rb.n_init_samples_ = (int)y.Num("n_init_samples", rb.n_init_samples_);
        rb.n_iterations_ = (int)y.Num("n_iterations", rb.n_iterations_);
        rb.random_seed_ = (int)y.Num("random_seed", rb.random_seed_);
        rb.verbose_level_ = (int)y.Num("verbose_level", rb.verbose_level_);
        rb.noise_ = y.Num("noise", rb.noise_);
        rb.n_iter_relearn_ = (int)y.Num("n_iteration_relearn", rb.n_iter_relearn_);
        rb.n_inner_evaluation_ = (int)y.Num("n_inner_evaluation", rb.n_inner_evaluation_);
        rb.lower_bound_ = y.List("lower_bound", rb.lower_bound_);
        rb.upper_bound_ = y.List("upper_bound", rb.upper_bound_);
        if (rb.lower_bound_.size() != rb.upper_bound_.size() || rb.lower_bound_.empty())
            throw Error(KFBO_EINVAL, path + ": lower_bound / upper_bound must have equal, non-zero length");
        return rb;
    }
};

class GpEi {
public:
    using Vec = std::vector<double>;

    GpEi(const ReadParaBayes &rb, uint64_t seed) : rb_(rb), dim_((int)rb.lower_bound_.size()), rng_(seed), ell_(dim_, 0.3) {}

    // Minimises f over the box; returns the best point found (in box coordinates).
    Vec Optimize(const std::function<double(const Vec &)> &f)
    {
        // ---- initial design: Latin hypercube
        const int n0 = std::max(rb_.n_init_samples_, 2);
        std::vector<std::vector<int>> perm(dim_, std::vector<int>(n0));
        for (int d = 0; d < dim_; ++d) {
            for (int i = 0; i < n0; ++i) perm[d][i] = i;
            std::shuffle(perm[d].begin(), perm[d].end(), rng_);
        }
        std::uniform_real_distribution<double> U(0.0, 1.0);
        for (int i = 0; i < n0; ++i) {
            Vec u(dim_);
            for (int d = 0; d < dim_; ++d) u[d] = (perm[d][i] + U(rng_)) / n0;
            Add(u, f(ToBox(u)));
        }
        // ---- sequential expected improvement
        for (int it = 0; it < rb_.n_iterations_; ++it) {
            if (it % std::max(rb_.n_iter_relearn_, 1) == 0) Relearn();
            Fit();
            const Vec u = MaximiseEi();
            Add(u, f(ToBox(u)));
        }
        Fit();   // the surrogate including the last sample (for Predict)
        return ToBox(X_[best_]);
    }

    double best_value() const { return y_[best_]; }
    size_t evaluations() const { return y_.size(); }
    const std::vector<Vec> &points() const { return X_; }
    const Vec &values() const { return y_; }

    // posterior mean / standard deviation at box point x (valid after Optimize or Fit)
    void Predict(const Vec &x, double *mu, double *sd) const
    {
        Vec u(dim_);
        for (int d = 0; d < dim_; ++d) u[d] = (x[d] - rb_.lower_bound_[d]) / (rb_.upper_bound_[d] - rb_.lower_bound_[d]);
        double m, s;
        PredictUnit(u, &m, &s);
        *mu = m * ysd_ + ymean_;
        *sd = s * ysd_;
    }

private:
    Vec ToBox(const Vec &u) const
    {
        Vec x(dim_);
        for (int d = 0; d < dim_; ++d) x[d] = rb_.lower_bound_[d] + u[d] * (rb_.upper_bound_[d] - rb_.lower_bound_[d]);
        return x;
    }
    void Add(const Vec &u, double y)
    {
        if (!std::isfinite(y)) y = 1e3;   // a NaN cost (log of a non-positive average) is a very bad sample
        X_.push_back(u);
        y_.push_back(y);
        if (y_.size() == 1 || y < y_[best_]) best_ = y_.size() - 1;
    }
    double Kernel(const Vec &a, const Vec &b, const Vec &ell) const   // Matern 3/2, ARD
    {
        double r2 = 0.0;
        for (int d = 0; d < dim_; ++d) { const double t = (a[d] - b[d]) / ell[d]; r2 += t * t; }
        const double r = std::sqrt(3.0 * r2);
        return (1.0 + r) * std::exp(-r);
    }
    // Cholesky of K + (noise + jitter) I for length-scales ell; returns false if not positive definite
    bool Factor(const Vec &ell, std::vector<double> *L, double jitter = 1e-10) const
    {
        const size_t n = X_.size();
        L->assign(n * n, 0.0);
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j <= i; ++j) (*L)[i * n + j] = Kernel(X_[i], X_[j], ell) + (i == j ? rb_.noise_ + jitter : 0.0);
        for (size_t j = 0; j < n; ++j) {
            double s = (*L)[j * n + j];
            for (size_t k = 0; k < j; ++k) s -= (*L)[j * n + k] * (*L)[j * n + k];
            if (!(s > 0.0)) return false;
            const double ljj = std::sqrt(s);
            (*L)[j * n + j] = ljj;
            for (size_t i = j + 1; i < n; ++i) {
                double t = (*L)[i * n + j];
                for (size_t k = 0; k < j; ++k) t -= (*L)[i * n + k] * (*L)[j * n + k];
                (*L)[i * n + j] = t / ljj;
            }
        }
        return true;
    }
    static void SolveLower(const std::vector<double> &L, size_t n, Vec *b)
    {
        for (size_t i = 0; i < n; ++i) {
            double t = (*b)[i];
            for (size_t k = 0; k < i; ++k) t -= L[i * n + k] * (*b)[k];
            (*b)[i] = t / L[i * n + i];
        }
    }
    static void SolveUpper(const std::vector<double> &L, size_t n, Vec *b)   // L^T x = b
    {
        for (size_t ii = n; ii-- > 0;) {
            double t = (*b)[ii];
            for (size_t k = ii + 1; k < n; ++k) t -= L[k * n + ii] * (*b)[k];
            (*b)[ii] = t / L[ii * n + ii];
        }
    }
    void Standardise()
    {
        const size_t n = y_.size();
        ymean_ = 0.0;
        for (double v : y_) ymean_ += v / n;
        double var = 0.0;
        for (double v : y_) var += (v - ymean_) * (v - ymean_) / std::max<size_t>(n - 1, 1);
        ysd_ = var > 1e-24 ? std::sqrt(var) : 1.0;
        ys_.resize(n);
        for (size_t i = 0; i < n; ++i) ys_[i] = (y_[i] - ymean_) / ysd_;
    }
    double NegLogMarginal(const Vec &ell)
    {
        std::vector<double> L;
        if (!Factor(ell, &L)) return std::numeric_limits<double>::infinity();
        const size_t n = X_.size();
        Vec a = ys_;
        SolveLower(L, n, &a);
        double q = 0.0, logdet = 0.0;
        for (size_t i = 0; i < n; ++i) { q += a[i] * a[i]; logdet += std::log(L[i * n + i]); }
        return 0.5 * q + logdet;
    }
    void Relearn()   // coordinate search over log length-scales
    {
        Standardise();
        double best = NegLogMarginal(ell_);
        for (int sweep = 0; sweep < 3; ++sweep)
            for (int d = 0; d < dim_; ++d)
                for (double f : {0.5, 0.7, 1.4, 2.0}) {
                    Vec e = ell_;
                    e[d] = std::min(std::max(e[d] * f, 0.02), 5.0);
                    const double v = NegLogMarginal(e);
                    if (v < best) { best = v; ell_ = e; }
                }
    }
    void Fit()
    {
        Standardise();
        // near-duplicate samples (the optimiser converging) make K numerically singular: escalate the diagonal jitter
        // instead of giving the whole loop up
        double jitter = 1e-10;
        while (!Factor(ell_, &L_, jitter)) {
            jitter *= 100.0;
            if (jitter > 1e-2) throw Error(KFBO_EINVAL, "GP covariance is not positive definite even with a jitter of 1e-2");
        }
        alpha_ = ys_;
        SolveLower(L_, X_.size(), &alpha_);
        SolveUpper(L_, X_.size(), &alpha_);
    }
    void PredictUnit(const Vec &u, double *mu, double *sd) const
    {
        const size_t n = X_.size();
        Vec k(n);
        double m = 0.0;
        for (size_t i = 0; i < n; ++i) { k[i] = Kernel(u, X_[i], ell_); m += k[i] * alpha_[i]; }
        SolveLower(L_, n, &k);
        double v = 1.0 + rb_.noise_;
        for (size_t i = 0; i < n; ++i) v -= k[i] * k[i];
        *mu = m;
        *sd = std::sqrt(std::max(v, 1e-18));
    }
    double Ei(const Vec &u) const
    {
        double mu, sd;
        PredictUnit(u, &mu, &sd);
        const double ymin = ys_[best_], z = (ymin - mu) / sd;
        const double Phi = 0.5 * std::erfc(-z / std::sqrt(2.0)), phi = std::exp(-0.5 * z * z) / std::sqrt(2.0 * M_PI);
        return (ymin - mu) * Phi + sd * phi;
    }
    Vec MaximiseEi()
    {
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::normal_distribution<double> N(0.0, 1.0);
        const int n_global = rb_.n_inner_evaluation_ * dim_;
        Vec best_u(dim_, 0.5);
        double best = -1.0;
        for (int i = 0; i < n_global; ++i) {
            Vec u(dim_);
            for (int d = 0; d < dim_; ++d) u[d] = U(rng_);
            const double e = Ei(u);
            if (e > best) { best = e; best_u = u; }
        }
        double step = 0.05;   // local refinement around the best candidate and around the incumbent
        for (int i = 0; i < 6 * std::max(rb_.n_inner_evaluation_ / 10, 10); ++i) {
            Vec u(dim_);
            const Vec &c = (i & 1) ? X_[best_] : best_u;
            for (int d = 0; d < dim_; ++d) u[d] = std::min(std::max(c[d] + step * N(rng_), 0.0), 1.0);
            const double e = Ei(u);
            if (e > best) { best = e; best_u = u; } else if ((i & 7) == 7) step = std::max(step * 0.8, 1e-3);
        }
        return best_u;
    }

    ReadParaBayes rb_;
    int dim_;
    std::mt19937_64 rng_;
    Vec ell_;
    std::vector<Vec> X_;
    Vec y_, ys_, alpha_;
    std::vector<double> L_;
    double ymean_ = 0.0, ysd_ = 1.0;
    size_t best_ = 0;
};

}  // namespace bo
}  // namespace kfbo
