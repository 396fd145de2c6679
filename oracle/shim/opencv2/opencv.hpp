// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product.
//
// Minimal stand-in for <opencv2/opencv.hpp> so that the UNMODIFIED reference
// headers under /root/reference/cpp/volumetric compile in a container without
// the OpenCV C++ SDK.  Only the surface those headers touch is provided:
//   cv::Mat {rows, cols, type(), channels(), empty(), at<T>(), ptr<T>(),
//            convertTo(), setTo()}  and the CV_* depth macros.
// (reference users: cpp/volumetric/image_utils.h:30-165,
//  cpp/volumetric/voxel_grid_carving.h:47-80, cpp/volumetric/voxel_block_grid.h:52)
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)

namespace cv {

class Mat {
  public:
    int rows = 0;
    int cols = 0;

    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
        owned_ = std::make_shared<std::vector<unsigned char>>(bytes(), 0);
        data_ = owned_->data();
    }
    // borrowed view over caller memory (like cv::Mat(rows, cols, type, void*))
    Mat(int r, int c, int type, void *data)
        : rows(r), cols(c), type_(type), data_(static_cast<unsigned char *>(data)) {}

    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }

    static size_t depth_size(int depth) {
        switch (depth) {
        case CV_8U:
        case CV_8S:
            return 1;
        case CV_16U:
        case CV_16S:
            return 2;
        case CV_32S:
        case CV_32F:
            return 4;
        case CV_64F:
            return 8;
        default:
            throw std::runtime_error("cv shim: bad depth");
        }
    }
    size_t elem_size() const { return depth_size(depth()) * channels(); }
    size_t step() const { return elem_size() * static_cast<size_t>(cols); }
    size_t bytes() const { return step() * static_cast<size_t>(rows); }

    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data_ + step() * r); }
    template <typename T> const T *ptr(int r = 0) const {
        return reinterpret_cast<const T *>(data_ + step() * r);
    }
    template <typename T> T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }

    template <typename S> void setTo(S value) {
        const size_t n = static_cast<size_t>(rows) * cols * channels();
        switch (depth()) {
        case CV_8U: fill<uint8_t>(n, value); break;
        case CV_8S: fill<int8_t>(n, value); break;
        case CV_16U: fill<uint16_t>(n, value); break;
        case CV_16S: fill<int16_t>(n, value); break;
        case CV_32S: fill<int32_t>(n, value); break;
        case CV_32F: fill<float>(n, value); break;
        case CV_64F: fill<double>(n, value); break;
        }
    }

    void convertTo(Mat &dst, int rtype) const {
        Mat out(rows, cols, CV_MAKETYPE(CV_MAT_DEPTH(rtype), channels()));
        const size_t n = static_cast<size_t>(rows) * cols * channels();
        for (size_t i = 0; i < n; ++i) {
            out.store(i, load(i));
        }
        dst = out;
    }

  private:
    template <typename T, typename S> void fill(size_t n, S value) {
        T *p = reinterpret_cast<T *>(data_);
        for (size_t i = 0; i < n; ++i) p[i] = static_cast<T>(value);
    }
    double load(size_t i) const {
        switch (depth()) {
        case CV_8U: return reinterpret_cast<const uint8_t *>(data_)[i];
        case CV_8S: return reinterpret_cast<const int8_t *>(data_)[i];
        case CV_16U: return reinterpret_cast<const uint16_t *>(data_)[i];
        case CV_16S: return reinterpret_cast<const int16_t *>(data_)[i];
        case CV_32S: return reinterpret_cast<const int32_t *>(data_)[i];
        case CV_32F: return reinterpret_cast<const float *>(data_)[i];
        default: return reinterpret_cast<const double *>(data_)[i];
        }
    }
    void store(size_t i, double v) {
        switch (depth()) {
        case CV_8U: reinterpret_cast<uint8_t *>(data_)[i] = static_cast<uint8_t>(v); break;
        case CV_8S: reinterpret_cast<int8_t *>(data_)[i] = static_cast<int8_t>(v); break;
        case CV_16U: reinterpret_cast<uint16_t *>(data_)[i] = static_cast<uint16_t>(v); break;
        case CV_16S: reinterpret_cast<int16_t *>(data_)[i] = static_cast<int16_t>(v); break;
        case CV_32S: reinterpret_cast<int32_t *>(data_)[i] = static_cast<int32_t>(v); break;
        case CV_32F: reinterpret_cast<float *>(data_)[i] = static_cast<float>(v); break;
        default: reinterpret_cast<double *>(data_)[i] = v; break;
        }
    }

    int type_ = 0;
    unsigned char *data_ = nullptr;
    std::shared_ptr<std::vector<unsigned char>> owned_;
};

} // namespace cv
