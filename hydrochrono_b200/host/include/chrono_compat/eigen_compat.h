// Minimal stand-ins for the few Eigen types that appear in HydroChrono's public signatures
// (Eigen::VectorXd / Vector3d / MatrixXd).  Eigen is not installable in this image; when the real Eigen is
// available compile with -DHYDROC_HAVE_EIGEN and these are not used.  Only what the hydro API needs is provided:
// element access, sizes, raw data.  No expression templates, no linear algebra.
#pragma once
#ifdef HYDROC_HAVE_EIGEN
#include <Eigen/Dense>
#else
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace Eigen {

class VectorXd {
  public:
    VectorXd() = default;
    explicit VectorXd(std::ptrdiff_t n) : v_(static_cast<size_t>(n), 0.0) {}
    VectorXd(std::initializer_list<double> l) : v_(l) {}
    explicit VectorXd(const std::vector<double>& v) : v_(v) {}
    std::ptrdiff_t size() const { return static_cast<std::ptrdiff_t>(v_.size()); }
    std::ptrdiff_t rows() const { return size(); }
    void resize(std::ptrdiff_t n) { v_.resize(static_cast<size_t>(n)); }
    void setZero() { v_.assign(v_.size(), 0.0); }
    double& operator[](std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
    double operator[](std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
    double& operator()(std::ptrdiff_t i) { return v_[static_cast<size_t>(i)]; }
    double operator()(std::ptrdiff_t i) const { return v_[static_cast<size_t>(i)]; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double* begin() { return v_.data(); }
    double* end() { return v_.data() + v_.size(); }
    const double* begin() const { return v_.data(); }
    const double* end() const { return v_.data() + v_.size(); }
    const std::vector<double>& std() const { return v_; }

  private:
    std::vector<double> v_;
};

class Vector3d {
  public:
    Vector3d() : d_{0, 0, 0} {}
    Vector3d(double x, double y, double z) : d_{x, y, z} {}
    double& x() { return d_[0]; }
    double& y() { return d_[1]; }
    double& z() { return d_[2]; }
    double x() const { return d_[0]; }
    double y() const { return d_[1]; }
    double z() const { return d_[2]; }
    double& operator[](int i) { return d_[i]; }
    double operator[](int i) const { return d_[i]; }
    Vector3d& operator+=(const Vector3d& o) { d_[0] += o.d_[0]; d_[1] += o.d_[1]; d_[2] += o.d_[2]; return *this; }

  private:
    double d_[3];
};
template <class T>
using Vector3 = Vector3d;

// Row/column indexed dense matrix (storage order is an implementation detail of the stand-in).
class MatrixXd {
  public:
    MatrixXd() = default;
    MatrixXd(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
    void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; v_.assign(static_cast<size_t>(r * c), 0.0); }
    void setZero() { v_.assign(v_.size(), 0.0); }
    void setZero(std::ptrdiff_t r, std::ptrdiff_t c) { resize(r, c); }
    std::ptrdiff_t rows() const { return r_; }
    std::ptrdiff_t cols() const { return c_; }
    std::ptrdiff_t size() const { return r_ * c_; }
    double& operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return v_[static_cast<size_t>(i * c_ + j)]; }
    double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return v_[static_cast<size_t>(i * c_ + j)]; }
    double* data() { return v_.data(); }               // row-major
    const double* data() const { return v_.data(); }

  private:
    std::ptrdiff_t r_ = 0, c_ = 0;
    std::vector<double> v_;
};

}  // namespace Eigen
#endif
