// cv_standin.h — the minimal subset of OpenCV's cv::Mat / cv::KeyPoint / cv::Point2f that the reference's matcher
// signatures mention (SfM/src/feature/feature_matching.h:33-63).  OpenCV's C++ headers are not in this image; when the
// shim is compiled inside MetricSfM define MSFM_USE_OPENCV and the real <opencv2/opencv.hpp> types are used instead.
#pragma once
#ifdef MSFM_USE_OPENCV
#include <opencv2/opencv.hpp>
#else
#include <cstddef>
#include <cstdint>
#include <vector>

#define CV_32FC1 5
#define CV_8UC1 0

namespace cv {
struct Point2f {
    float x = 0.f, y = 0.f;
};
struct KeyPoint {
    Point2f pt;
    float size = 0.f, angle = -1.f, response = 0.f;
    int octave = 0, class_id = -1;
};
// Row-major matrix view/owner with the members the matcher touches: rows, cols, data, step, type(), ptr<T>(row).
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;  // bytes per row
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
        step = (size_t)c * (type == CV_32FC1 ? 4 : 1);
        store_.resize((size_t)r * step);
        data = store_.data();
    }
    Mat(int r, int c, int type, void *external, size_t step_bytes = 0) : rows(r), cols(c), data((unsigned char *)external), type_(type) {
        step = step_bytes ? step_bytes : (size_t)c * (type == CV_32FC1 ? 4 : 1);
    }
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    template <typename T>
    T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
    template <typename T>
    const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }

private:
    int type_ = CV_32FC1;
    std::vector<unsigned char> store_;
};
}  // namespace cv
#endif
