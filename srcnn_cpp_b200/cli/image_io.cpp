// image_io.cpp -- PNG (zlib) and PPM/PGM codecs for the CLI.  See image_io.h.
#include "image_io.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <new>

bool jpeg_decode(const std::vector<uint8_t>& file, ImageBGR* out, std::string* err);   // jpeg_io.cpp (nvJPEG)
bool jpeg_encode(const ImageBGR& img, std::vector<uint8_t>* file, std::string* err);

namespace {

bool read_file(const std::string& path, std::vector<uint8_t>* buf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf->resize((size_t)n);
    size_t got = n ? fread(buf->data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool png_decode(const std::vector<uint8_t>& file, ImageBGR* out, std::string* err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { *err = "not a PNG"; return false; }
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) { *err = "truncated PNG"; return false; }
        const uint8_t* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (w == 0 || h == 0) { *err = "PNG without IHDR"; return false; }
    // IHDR is untrusted: bound the geometry before any size arithmetic (cv::imread's own limit is 2^30 pixels)
    if (w > (1u << 20) || h > (1u << 20) || (uint64_t)w * h > (1ull << 30)) { *err = "PNG dimensions out of range"; return false; }
    if (interlace) { *err = "interlaced (Adam7) PNG not supported"; return false; }
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) { *err = "PNG colour type not supported"; return false; }
    const bool sub_byte = depth == 1 || depth == 2 || depth == 4;       // grey or palette only (PNG spec)
    if (!(depth == 8 || depth == 16 || (sub_byte && (ctype == 0 || ctype == 3))) || (ctype == 3 && depth == 16)) {
        *err = "PNG bit depth not supported for this colour type";
        return false;
    }
    const size_t bpp = std::max<size_t>(1, (size_t)ch * depth / 8);     // filter distance in bytes
    const size_t stride = ((size_t)ch * depth * w + 7) / 8;             // packed scanline
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) { *err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t* src = &raw[y * (stride + 1) + 1];
        uint8_t* cur = &img[y * stride];
        const uint8_t* up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0, v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: *err = "bad PNG filter"; return false;
            }
            cur[x] = (uint8_t)v;
        }
    }
    out->w = (int)w; out->h = (int)h;
    out->px.resize((size_t)w * h * 3);
    if (sub_byte) {   // 1/2/4-bit grey (scaled to 0..255) or palette indices, most significant bits first
        const int per_byte = 8 / depth, mask = (1 << depth) - 1, grey_mul = 255 / mask;
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const int v = (img[y * stride + x / per_byte] >> ((per_byte - 1 - x % per_byte) * depth)) & mask;
                uint8_t r, g, b;
                if (ctype == 3) {
                    const size_t k = (size_t)v * 3;
                    if (k + 2 < plte.size()) { r = plte[k]; g = plte[k + 1]; b = plte[k + 2]; } else { r = g = b = 0; }
                } else { r = g = b = (uint8_t)(v * grey_mul); }
                uint8_t* q = &out->px[((size_t)y * w + x) * 3];
                q[0] = b; q[1] = g; q[2] = r;
            }
        return true;
    }
    const size_t step = depth / 8;  // 16-bit samples: keep the high byte (like cv::imread's 8-bit conversion)
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const uint8_t* p = &img[i * bpp];
        uint8_t r, g, b;
        if (ctype == 0 || ctype == 4) { r = g = b = p[0]; }
        else if (ctype == 3) {
            const size_t k = (size_t)p[0] * 3;
            if (k + 2 < plte.size()) { r = plte[k]; g = plte[k + 1]; b = plte[k + 2]; } else { r = g = b = 0; }
        } else { r = p[0]; g = p[step]; b = p[2 * step]; }
        out->px[3 * i] = b; out->px[3 * i + 1] = g; out->px[3 * i + 2] = r;  // alpha is dropped, like IMREAD_COLOR
    }
    return true;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

void png_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
    put_be32(out, (uint32_t)len);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    uint32_t crc = crc32(0, &out[start], (uInt)(len + 4));
    put_be32(out, crc);
}

bool png_encode(const ImageBGR& img, std::vector<uint8_t>* file) {
    const size_t stride = (size_t)img.w * 3;
    std::vector<uint8_t> raw((stride + 1) * img.h);
    for (int y = 0; y < img.h; y++) {
        uint8_t* row = &raw[(size_t)y * (stride + 1)];
        row[0] = 0;  // filter: none (deterministic and simple; zlib does the rest)
        const uint8_t* s = &img.px[(size_t)y * stride];
        for (int x = 0; x < img.w; x++) { row[1 + 3 * x] = s[3 * x + 2]; row[2 + 3 * x] = s[3 * x + 1]; row[3 + 3 * x] = s[3 * x]; }
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    file->assign(sig, sig + 8);
    uint8_t ihdr[13];
    ihdr[0] = img.w >> 24; ihdr[1] = img.w >> 16; ihdr[2] = img.w >> 8; ihdr[3] = img.w;
    ihdr[4] = img.h >> 24; ihdr[5] = img.h >> 16; ihdr[6] = img.h >> 8; ihdr[7] = img.h;
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    png_chunk(*file, "IHDR", ihdr, 13);
    png_chunk(*file, "IDAT", comp.data(), clen);
    png_chunk(*file, "IEND", nullptr, 0);
    return true;
}

bool pnm_decode(const std::vector<uint8_t>& f, ImageBGR* out, std::string* err) {
    if (f.size() < 2 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) { *err = "not a binary PPM/PGM"; return false; }
    const int ch = f[1] == '6' ? 3 : 1;
    size_t pos = 2;
    long vals[3];
    for (int k = 0; k < 3; k++) {
        for (;;) {
            while (pos < f.size() && isspace(f[pos])) pos++;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; continue; }
            break;
        }
        long v = 0; bool any = false;
        while (pos < f.size() && isdigit(f[pos])) { v = std::min(v * 10 + (f[pos] - '0'), 1L << 40); pos++; any = true; }
        if (!any) { *err = "bad PNM header"; return false; }
        vals[k] = v;
    }
    pos++;  // single whitespace after maxval
    if (vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0) { *err = "only 8-bit PNM supported"; return false; }
    if (vals[0] > (1L << 20) || vals[1] > (1L << 20) || vals[0] * vals[1] > (1L << 30)) { *err = "PNM dimensions out of range"; return false; }
    const size_t n = (size_t)vals[0] * vals[1];
    if (pos > f.size() || n * ch > f.size() - pos) { *err = "truncated PNM"; return false; }
    out->w = (int)vals[0]; out->h = (int)vals[1]; out->px.resize(n * 3);
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = &f[pos + i * ch];
        if (ch == 3) { out->px[3 * i] = p[2]; out->px[3 * i + 1] = p[1]; out->px[3 * i + 2] = p[0]; }
        else { out->px[3 * i] = out->px[3 * i + 1] = out->px[3 * i + 2] = p[0]; }
    }
    return true;
}

std::string lower_ext(const std::string& path) {
    size_t dot = path.find_last_of('.');
    std::string e = dot == std::string::npos ? "" : path.substr(dot);
    for (auto& c : e) c = (char)tolower(c);
    return e;
}

}  // namespace

// Known gap against cv::imread (src/srcnn.cpp:462): Adam7-interlaced PNGs, ASCII PNM, BMP/TIFF/WebP are reported as a load
// failure (the reference CLI's "- load failure" path, exit code -1).
bool image_read(const std::string& path, ImageBGR* out, std::string* err) {
    try {   // a crafted header must end in the load-failure path, not in an uncaught bad_alloc inside the worker thread
        std::vector<uint8_t> f;
        if (!read_file(path, &f)) { *err = "cannot read file"; return false; }
        if (f.size() >= 8 && f[0] == 0x89 && f[1] == 'P') return png_decode(f, out, err);
        if (f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) return pnm_decode(f, out, err);
        if (f.size() >= 3 && f[0] == 0xff && f[1] == 0xd8 && f[2] == 0xff) return jpeg_decode(f, out, err);
        *err = "unsupported image format (PNG, JPEG and binary PPM/PGM are supported)";
    } catch (const std::bad_alloc&) {
        *err = "out of memory while decoding";
    } catch (const std::exception& e) {
        *err = e.what();
    }
    return false;
}

bool image_write(const std::string& path, const ImageBGR& img, std::string* err) {
    std::vector<uint8_t> file;
    const std::string e = lower_ext(path);
    if (e == ".ppm" || e == ".pnm") {
        char hdr[64];
        int n = snprintf(hdr, sizeof(hdr), "P6\n%d %d\n255\n", img.w, img.h);
        file.assign(hdr, hdr + n);
        for (size_t i = 0; i < (size_t)img.w * img.h; i++) { file.push_back(img.px[3 * i + 2]); file.push_back(img.px[3 * i + 1]); file.push_back(img.px[3 * i]); }
    } else if (e == ".png" || e.empty()) {
        if (!png_encode(img, &file)) { *err = "PNG encode failed"; return false; }
    } else if (e == ".jpg" || e == ".jpeg") {
        if (!jpeg_encode(img, &file, err)) return false;
    } else {
        *err = "unsupported output format '" + e + "' (use .png, .jpg or .ppm)";
        return false;
    }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { *err = "cannot open output file"; return false; }
    size_t put = fwrite(file.data(), 1, file.size(), f);
    fclose(f);
    if (put != file.size()) { *err = "short write"; return false; }
    return true;
}
