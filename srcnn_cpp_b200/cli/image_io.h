// image_io.h -- minimal image file I/O for the bin/srcnn drop-in (the reference uses cv::imread /
// cv::imwrite, src/srcnn.cpp:462,670; OpenCV C++ is not available here).  PNG over zlib and binary
// PPM/PGM.  Images are returned / taken as 8-bit BGR, HWC, like cv::imread(IMREAD_COLOR).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct ImageBGR {
    int w = 0, h = 0;
    std::vector<uint8_t> px;  // h * w * 3, B,G,R
    bool empty() const { return w <= 0 || h <= 0 || px.empty(); }
};

bool image_read(const std::string& path, ImageBGR* out, std::string* err);
bool image_write(const std::string& path, const ImageBGR& img, std::string* err);
