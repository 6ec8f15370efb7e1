// Stand-in so the reference's bit_decompose.hh / mat_decompose.hh parse without nanobind
// (absent from this image).  Only declarations; the oracle never calls into Python glue.
#pragma once
namespace nanobind {
struct tuple {};
template <class... A> struct ndarray {};
namespace literals {}
} // namespace nanobind
