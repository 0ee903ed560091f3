// dm.hpp -- minimal dense value types standing in for casadi::DM / casadi::Dict / casadi::Function on the
// kite_model API surface (CasADi is a third-party dependency of the reference and is not used here).
// Only what the callers of the hot path need: construction, element access, a few algebra helpers.
#pragma once
#include <cmath>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace openkite {

/** Dense matrix, column-major like casadi::DM's dense storage; a vector is n x 1. */
class DM {
public:
    DM() : r_(0), c_(0) {}
    DM(double v) : r_(1), c_(1), d_(1, v) {}
    DM(std::initializer_list<double> v) : r_((int)v.size()), c_(1), d_(v) {}
    explicit DM(const std::vector<double>& v) : r_((int)v.size()), c_(1), d_(v) {}
    DM(int rows, int cols, double fill = 0.0) : r_(rows), c_(cols), d_((size_t)rows * cols, fill) {}

    static DM zeros(int r, int c = 1) { return DM(r, c, 0.0); }
    static DM ones(int r, int c = 1) { return DM(r, c, 1.0); }
    static DM eye(int n) { DM m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
    static DM diag(const DM& v) { int n = v.numel(); DM m(n, n); for (int i = 0; i < n; ++i) m(i, i) = v.d_[i]; return m; }
    static DM vertcat(const std::vector<DM>& parts) {
        std::vector<double> out;
        for (const auto& p : parts) { if (p.c_ > 1) throw std::invalid_argument("DM::vertcat: vectors only"); out.insert(out.end(), p.d_.begin(), p.d_.end()); }
        return DM(out);
    }
    static DM mtimes(const DM& a, const DM& b) {
        if (a.c_ != b.r_) throw std::invalid_argument("DM::mtimes: dimension mismatch");
        DM m(a.r_, b.c_);
        for (int j = 0; j < b.c_; ++j) for (int k = 0; k < a.c_; ++k) { double bkj = b(k, j); for (int i = 0; i < a.r_; ++i) m(i, j) += a(i, k) * bkj; }
        return m;
    }
    DM T() const { DM m(c_, r_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) m(j, i) = (*this)(i, j); return m; }

    int size1() const { return r_; }
    int size2() const { return c_; }
    int numel() const { return r_ * c_; }
    bool is_empty() const { return d_.empty(); }
    double& operator()(int i, int j = 0) { return d_[(size_t)j * r_ + i]; }
    double operator()(int i, int j = 0) const { return d_[(size_t)j * r_ + i]; }
    double& operator[](int i) { return d_[i]; }
    double operator[](int i) const { return d_[i]; }
    const std::vector<double>& nonzeros() const { return d_; }
    std::vector<double>& nonzeros() { return d_; }
    const double* ptr() const { return d_.data(); }
    double* ptr() { return d_.data(); }
    /** row-major copy (the engine's matrix component order is i*cols + j) */
    std::vector<double> row_major() const { std::vector<double> o((size_t)r_ * c_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) o[(size_t)i * c_ + j] = (*this)(i, j); return o; }
    static DM from_row_major(const double* p, int r, int c) { DM m(r, c); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = p[(size_t)i * c + j]; return m; }

    DM operator+(const DM& o) const { return zip(o, [](double a, double b) { return a + b; }); }
    DM operator-(const DM& o) const { return zip(o, [](double a, double b) { return a - b; }); }
    DM operator*(double s) const { DM m = *this; for (auto& v : m.d_) v *= s; return m; }
    friend DM operator*(double s, const DM& a) { return a * s; }
    static double norm_inf(const DM& a) { double m = 0; for (double v : a.d_) m = std::fmax(m, std::fabs(v)); return m; }

private:
    template <class F> DM zip(const DM& o, F f) const {
        if (o.numel() == 1 && numel() != 1) { DM m = *this; for (auto& v : m.d_) v = f(v, o.d_[0]); return m; }
        if (r_ != o.r_ || c_ != o.c_) throw std::invalid_argument("DM: shape mismatch");
        DM m = *this; for (size_t i = 0; i < d_.size(); ++i) m.d_[i] = f(d_[i], o.d_[i]); return m;
    }
    int r_, c_;
    std::vector<double> d_;
};
typedef std::vector<DM> DMVector;

/** Option dictionary standing in for casadi::Dict (numeric options only: method, tf, tol, max_iter, ...). */
typedef std::map<std::string, double> Dict;

/** Named numeric function handle standing in for casadi::Function: evaluates on the GPU engine (B = 1). */
class Function {
public:
    typedef std::function<DMVector(const DMVector&)> Eval;
    Function() : n_in_(0), n_out_(0) {}
    Function(const std::string& name, std::vector<int> in_sizes, std::vector<int> out_sizes, Eval f, std::shared_ptr<void> owner = nullptr)
        : name_(name), in_(std::move(in_sizes)), out_(std::move(out_sizes)), n_in_((int)in_.size()), n_out_((int)out_.size()), f_(std::move(f)), owner_(std::move(owner)) {}
    const std::string& name() const { return name_; }
    bool is_null() const { return !f_; }
    int n_in() const { return n_in_; }
    int n_out() const { return n_out_; }
    /** total number of input / output elements (casadi::Function::nnz_in / nnz_out) */
    int nnz_in() const { int s = 0; for (int v : in_) s += v; return s; }
    int nnz_out() const { int s = 0; for (int v : out_) s += v; return s; }
    int size_in(int i) const { return in_.at(i); }
    DMVector operator()(const DMVector& args) const {
        if (!f_) throw std::runtime_error("Function '" + name_ + "' is null");
        if ((int)args.size() != n_in_) throw std::invalid_argument("Function '" + name_ + "': wrong number of inputs");
        for (int i = 0; i < n_in_; ++i) if (args[i].numel() != in_[i]) throw std::invalid_argument("Function '" + name_ + "': wrong input size");
        return f_(args);
    }
    /** the engine context this function evaluates on (opaque; see openkite::KiteContext) */
    const std::shared_ptr<void>& owner() const { return owner_; }

private:
    std::string name_;
    std::vector<int> in_, out_;
    int n_in_, n_out_;
    Eval f_;
    std::shared_ptr<void> owner_;
};

}  // namespace openkite
