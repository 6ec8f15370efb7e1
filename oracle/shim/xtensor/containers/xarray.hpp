// Stand-in for xtensor's xt::xarray<T>, TEST INFRASTRUCTURE ONLY.
//
// The reference's CMVM core (api.cc, cmvm_core.cc, state_opr.cc, indexers.cc,
// types.hh under /root/reference/src/da4ml/_binary/cmvm/) only needs a dense
// row-major nd-container: shape()/shape(i)/dimension()/size(), element access
// (i), (i,j), (i,j,k), begin()/end(), copy, and `xt::view(a, i) = scalar`
// (state_opr.cc:95).  xtensor itself is absent from this image (it is a meson
// wrap fetched from the network), so this header lets those four TUs compile
// *unmodified and in place*.  No algorithmic content lives here.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <stdexcept>
#include <vector>

namespace xt {

template <class T> class xarray {
  public:
    using value_type = T;
    using shape_type = std::vector<size_t>;

    xarray() = default;
    explicit xarray(const shape_type &shape, T fill = T()) { resize(shape, fill); }
    xarray(std::initializer_list<T> vals) : shape_{vals.size()}, data_(vals) {}

    void resize(const shape_type &shape, T fill = T()) {
        shape_ = shape;
        size_t n = 1;
        for (size_t s : shape_)
            n *= s;
        data_.assign(n, fill);
    }

    const shape_type &shape() const { return shape_; }
    size_t shape(size_t i) const { return shape_[i]; }
    size_t dimension() const { return shape_.size(); }
    size_t size() const { return data_.size(); }

    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }

    auto begin() { return data_.begin(); }
    auto end() { return data_.end(); }
    auto begin() const { return data_.begin(); }
    auto end() const { return data_.end(); }

    T &operator()(size_t i) { return data_[i]; }
    const T &operator()(size_t i) const { return data_[i]; }
    T &operator()(size_t i, size_t j) { return data_[i * shape_[1] + j]; }
    const T &operator()(size_t i, size_t j) const { return data_[i * shape_[1] + j]; }
    T &operator()(size_t i, size_t j, size_t k) {
        return data_[(i * shape_[1] + j) * shape_[2] + k];
    }
    const T &operator()(size_t i, size_t j, size_t k) const {
        return data_[(i * shape_[1] + j) * shape_[2] + k];
    }

    void fill(T v) { std::fill(data_.begin(), data_.end(), v); }

  private:
    shape_type shape_;
    std::vector<T> data_;
};

// `xt::view(arr, i) = scalar`  — fills the hyper-row i of the leading axis.
template <class T> struct row_view_ {
    xarray<T> &a;
    size_t i;
    row_view_ &operator=(T v) {
        size_t stride = a.shape(0) ? a.size() / a.shape(0) : 0;
        std::fill(a.data() + i * stride, a.data() + (i + 1) * stride, v);
        return *this;
    }
};
template <class T> row_view_<T> view(xarray<T> &a, size_t i) { return {a, i}; }

// Names that the reference's bit_decompose.hh mentions inside *uninstantiated*
// templates (_shift_amount, _center, bit_decompose.hh:21-34).  They only have to
// be declared for two-phase lookup; nothing in the oracle build instantiates them.
struct all_tag_ {};
struct newaxis_tag_ {};
inline all_tag_ all() { return {}; }
inline newaxis_tag_ newaxis() { return {}; }
struct never_ {
    template <class... A> never_ operator()(A &&...) const { return {}; }
};
template <class F> never_ vectorize(F &&) { return {}; }
template <class... A> never_ amin(A &&...) { return {}; }
template <class... A> never_ pow(A &&...) { return {}; }
template <class A, class B, class... R> never_ view(A &&, B &&, R &&...) { return {}; }
template <class A> never_ operator*(const A &, const never_ &) { return {}; }
template <class A> never_ operator-(const A &) requires requires(A a) { a.shape(); } { return {}; }

} // namespace xt
