// Test-infrastructure stub (NOT product code): the minimum of xt::pytensor that the
// reference's __PYTHONCC__-guarded constructors touch (shape/size/data/[]/(i,j)/begin/end),
// so that the unmodified reference headers under /root/reference/include can be
// instantiated from raw arrays without xtensor-python (absent in this image).
#pragma once
#include <array>
#include <vector>
#include <cstddef>

namespace xt {

template<typename T, std::size_t dim>
struct pytensor {
    std::array<long int, dim> shape_;
    std::vector<T> buf;

    pytensor() { shape_.fill(0); }
    explicit pytensor(const std::array<long int, dim>& shape) : shape_(shape) {
        std::size_t n = 1; for(auto s : shape) n *= (std::size_t)s;
        buf.resize(n);
    }
    pytensor(const T* src, const std::array<long int, dim>& shape) : pytensor(shape) {
        for(std::size_t i = 0; i < buf.size(); i++) buf[i] = src[i];
    }
    const std::array<long int, dim>& shape() const { return shape_; }
    std::size_t size() const { return buf.size(); }
    T* data() { return buf.data(); }
    const T* data() const { return buf.data(); }
    T& operator[](std::size_t i) { return buf[i]; }
    const T& operator[](std::size_t i) const { return buf[i]; }
    T& operator()(std::size_t i, std::size_t j) { return buf[i * shape_[1] + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return buf[i * shape_[1] + j]; }
    typename std::vector<T>::const_iterator begin() const { return buf.begin(); }
    typename std::vector<T>::const_iterator end() const { return buf.end(); }
};

inline void import_numpy() {}

} // namespace xt
