// cli.cpp -- `b2resample`: VkResample's command-line surface on top of the C ABI.
//
// Mirrors main() / launchResample() of the reference (VkResample.cpp:1782-1977 and :1280-1780):
// same flags (-h -devices -d -u -p -s -n -i -o -ifolder -ofolder -numfiles -numthreads), same
// per-thread frame loop (thread t handles files f*numThreads + t + 1, "%s/%06d.png" names,
// :1357,1624-1629,1750), same host-side pixel conversions (u8/255 in, truncating 255*v out,
// :1636-1685, :1708-1748) and the same stdout lines.  The Vulkan plumbing, VkFFT plans and shader
// dispatches are replaced by b2r_plan_create / upload / execute / download.  PNG I/O is a small
// zlib-based codec (the reference vendors stb_image; the image has zlib only).
// Additions: -gpus N spreads the worker threads over N CUDA devices (thread t -> device
// (d + t) % N); everything else behaves like the reference.
#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "b2resample.h"

// ------------------------------------------------------------------------------------------- PNG
namespace png {

static uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3]; }
static void put32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Decodes a non-interlaced PNG to 8-bit RGB (what stbi_load(..., 3) hands the reference:
// grey is replicated, alpha is dropped, 16-bit samples keep their high byte).
static bool load_rgb(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) { *err = "cannot open file"; return false; }
    std::vector<unsigned char> buf;
    unsigned char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 33 || memcmp(buf.data(), sig, 8)) { *err = "not a PNG file"; return false; }
    size_t pos = 8;
    int width = 0, height = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte;
    while (pos + 12 <= buf.size()) {
        uint32_t len = be32(&buf[pos]);
        const unsigned char* type = &buf[pos + 4];
        const unsigned char* data = &buf[pos + 8];
        if (pos + 12 + len > buf.size()) { *err = "truncated PNG"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            width = (int)be32(data); height = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (width <= 0 || height <= 0) { *err = "missing IHDR"; return false; }
    if (interlace) { *err = "interlaced PNG is not supported"; return false; }
    int channels = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
    if (!channels || (depth != 8 && depth != 16 && !(depth < 8 && (ctype == 0 || ctype == 3)))) {
        *err = "unsupported PNG colour type / bit depth"; return false;
    }
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t stride = ((size_t)width * bpp_bits + 7) / 8;
    const size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    std::vector<unsigned char> raw((stride + 1) * height);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) {
        *err = "zlib inflate failed"; return false;
    }
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    rgb->assign((size_t)width * height * 3, 0);
    for (int y = 0; y < height; ++y) {
        const unsigned char* row = &raw[(stride + 1) * y];
        const int ft = row[0];
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = row[1 + i];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) >> 1; break;
                case 4: x += paeth(a, b, c); break;
                default: *err = "bad PNG filter"; return false;
            }
            cur[i] = (unsigned char)x;
        }
        unsigned char* out = &(*rgb)[(size_t)y * width * 3];
        for (int x = 0; x < width; ++x) {
            unsigned char s[4] = {0, 0, 0, 255};
            if (depth == 8) for (int c = 0; c < channels; ++c) s[c] = cur[(size_t)x * channels + c];
            else if (depth == 16) for (int c = 0; c < channels; ++c) s[c] = cur[((size_t)x * channels + c) * 2];
            else {  // 1/2/4-bit grey or palette index
                int per = 8 / depth, v = (cur[x / per] >> (8 - depth * (x % per + 1))) & ((1 << depth) - 1);
                s[0] = (ctype == 3) ? (unsigned char)v : (unsigned char)(v * 255 / ((1 << depth) - 1));
            }
            if (ctype == 3) {
                size_t k = (size_t)s[0] * 3;
                if (k + 2 < plte.size()) { out[3 * x] = plte[k]; out[3 * x + 1] = plte[k + 1]; out[3 * x + 2] = plte[k + 2]; }
            } else if (ctype == 0 || ctype == 4) {
                out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = s[0];
            } else {
                out[3 * x] = s[0]; out[3 * x + 1] = s[1]; out[3 * x + 2] = s[2];
            }
        }
        prev.swap(cur);
    }
    *w = width; *h = height;
    return true;
}

// 8-bit RGB encoder with per-row adaptive filtering (minimum sum of absolute differences)
static bool write_rgb(const char* path, const unsigned char* rgb, int w, int h) {
    const size_t stride = (size_t)w * 3;
    std::vector<unsigned char> raw((stride + 1) * h), cand(stride), best(stride);
    std::vector<unsigned char> zero(stride, 0);
    for (int y = 0; y < h; ++y) {
        const unsigned char* cur = rgb + stride * y;
        const unsigned char* prev = y ? rgb + stride * (y - 1) : zero.data();
        long best_cost = -1; int best_ft = 0;
        for (int ft = 0; ft < 5; ++ft) {
            long cost = 0;
            for (size_t i = 0; i < stride; ++i) {
                int a = i >= 3 ? cur[i - 3] : 0, b = prev[i], c = i >= 3 ? prev[i - 3] : 0, x = cur[i];
                int v = ft == 0 ? x : ft == 1 ? x - a : ft == 2 ? x - b : ft == 3 ? x - ((a + b) >> 1) : x - paeth(a, b, c);
                cand[i] = (unsigned char)v;
                cost += std::abs((int)(signed char)cand[i]);
            }
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_ft = ft; best.swap(cand); }
        }
        raw[(stride + 1) * y] = (unsigned char)best_ft;
        memcpy(&raw[(stride + 1) * y + 1], best.data(), stride);
    }
    uLongf zlen = compressBound(raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), raw.size(), 6) != Z_OK) return false;
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char* type, const unsigned char* data, size_t len) {
        put32(out, (uint32_t)len);
        size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data, data + len);
        put32(out, (uint32_t)crc32(0, &out[start], (uInt)(len + 4)));
    };
    unsigned char ihdr[13];
    std::vector<unsigned char> t;
    put32(t, w); put32(t, h);
    memcpy(ihdr, t.data(), 8);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), zlen);
    chunk("IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}
}  // namespace png

// ------------------------------------------------------------------------------------------- CLI
struct Config {  // VkResampleConfiguration, VkResample.cpp:45-59
    uint32_t device_id = 0, upload_files = 0, num_iter = 1, precision = 0, num_threads = 1, thread_id = 0;
    uint32_t num_files = 1, gpus = 1, c2c = 0, fast = 0;
    float upscale = 1.0f, sharpen = 0.2f;
    const char* input = nullptr;
    const char* output = nullptr;
    const char* ifolder = nullptr;
    const char* ofolder = nullptr;
};

static bool find_flag(char** start, char** end, const std::string& flag) {      // VkResample.cpp:1782-1785
    for (char** p = start; p != end; ++p) if (flag == *p) return true;
    return false;
}
static char* flag_value(char** start, char** end, const std::string& flag) {    // VkResample.cpp:1786-1794
    for (char** p = start; p != end; ++p) if (flag == *p) return (p + 1 != end) ? *(p + 1) : nullptr;
    return nullptr;
}

// launchResample, VkResample.cpp:1280-1780
static int launch_resample(Config cfg) {
    const int ndev = b2r_device_count();
    if (ndev < 1) { printf("No CUDA device found: %s\n", b2r_last_error()); return -1; }
    const int device = (int)((cfg.device_id + (cfg.gpus > 1 ? cfg.thread_id % cfg.gpus : 0)) % (uint32_t)ndev);
    if (cfg.thread_id == 0) printf("VkResample - FFT based upscaling\n");
    char name[512];
    if (cfg.upload_files) snprintf(name, sizeof name, "%s/%06d.png", cfg.ifolder, cfg.thread_id + 1);
    else snprintf(name, sizeof name, "%s", cfg.input);
    std::vector<unsigned char> rgb;
    int w = 0, h = 0;
    std::string err;
    if (!png::load_rgb(name, &rgb, &w, &h, &err)) { printf("Image not found\n"); return 5; /* VK_INCOMPLETE */ }

    b2r_plan* plan = nullptr;
    int rc = b2r_plan_create(&plan, device, (uint32_t)w, (uint32_t)h, cfg.upscale, cfg.precision, cfg.sharpen,
                             (cfg.c2c ? B2R_FLAG_C2C_PARITY : B2R_FLAG_NONE) | (cfg.fast ? B2R_FLAG_FAST_SHARPEN : B2R_FLAG_NONE));
    if (rc) { printf("Plan creation failed, error code: %d (%s)\n", rc, b2r_last_error()); return rc; }
    b2r_plan_info info;
    b2r_plan_get_info(plan, &info);
    if (cfg.thread_id == 0)
        printf("VRAM per thread: %d MB Total: %d MB\n", (int)(info.device_bytes >> 20), (int)(cfg.num_threads * (info.device_bytes >> 20)));
    const size_t out_plane = (size_t)info.up_w * info.up_h;
    std::vector<unsigned char> png_out(out_plane * 3);

    uint32_t local_files = 1;
    if (cfg.upload_files) {  // VkResample.cpp:1622-1626
        local_files = (uint32_t)std::ceil(cfg.num_files / (float)cfg.num_threads);
        if ((local_files - 1) * cfg.num_threads + cfg.thread_id > cfg.num_files - 1) local_files--;
    }
    for (uint32_t f = 0; f < local_files; ++f) {
        if (f > 0) {
            snprintf(name, sizeof name, "%s/%06d.png", cfg.ifolder, f * cfg.num_threads + cfg.thread_id + 1);
            int w2, h2;
            if (!png::load_rgb(name, &rgb, &w2, &h2, &err)) { printf("Image not found\n"); return 5; }
            if (w2 != w || h2 != h) { printf("Image %s has a different size\n", name); return -1; }
        }
        // u8 HWC -> planar [0,1] (VkResample.cpp:1636-1685) runs on the GPU: b2r_upload_u8 ships the
        // interleaved bytes and converts there (same values: (float)((double)u8/255.0), or RN to half)
        if ((rc = b2r_upload_u8(plan, rgb.data()))) { printf("upload failed: %s\n", b2r_last_error()); return rc; }
        double ms = 0.0;
        if ((rc = b2r_execute(plan, cfg.num_iter, &ms))) { printf("execute failed: %s\n", b2r_last_error()); return rc; }
        if (!cfg.upload_files)
            printf("VkResample %0.1fx upscale: %dx%d to %dx%d Time: %0.3f ms\n", cfg.upscale, w, h, (int)info.up_w, (int)info.up_h, ms);
        // planar -> u8 HWC with the reference's truncating cast (:1715), also on the GPU
        if ((rc = b2r_download_u8(plan, png_out.data()))) { printf("download failed: %s\n", b2r_last_error()); return rc; }
        char oname[512];
        if (cfg.upload_files) snprintf(oname, sizeof oname, "%s/%06d.png", cfg.ofolder, f * cfg.num_threads + cfg.thread_id + 1);
        else if (cfg.output) snprintf(oname, sizeof oname, "%s", cfg.output);
        else snprintf(oname, sizeof oname, "%d_%d_upscaled.png", w, (int)info.up_w);  // :1706
        if (!png::write_rgb(oname, png_out.data(), (int)info.up_w, (int)info.up_h)) printf("cannot write %s\n", oname);
    }
    char dev_name[256] = "";
    b2r_device_name(device, dev_name, sizeof dev_name);
    printf("Thread %d finished. Device name: %s API:%s\n", cfg.thread_id, dev_name, b2r_version());
    b2r_plan_destroy(plan);
    return 0;
}

int main(int argc, char* argv[]) {
    Config cfg;
    cfg.upscale = 1.0f;   // reference defaults, VkResample.cpp:1798-1804
    if (find_flag(argv, argv + argc, "-h")) {
        printf("VkResample v1.0.2 (16-01-2021) CLI, served by b2resample (CUDA sm_100a). Based on the VkResample command line:\n"
               "\t-h: print help\n"
               "\t-devices: print the list of available GPU devices\n"
               "\t-d X: select GPU device (default 0)\n"
               "\t-u X: specify upscale factor (float, X>=1)\n"
               "\t-p X: specify precision (0 - single (default), 1 - double, 2 - half storage)\n"
               "\t-s X: specify sharpen constant (default 0.2)\n"
               "\t-n X: specify how many times to perform upscale. This removes dispatch overhead and will show the real application performance (default 1)\n"
               "\t-i NAME: specify input png file path\n"
               "\t-o NAME: specify output png file path (default X_Y_upscaled.png)\n"
               "\t-ifolder NAME: input folder; files are read as NAME/000001.png, NAME/000002.png, ...\n"
               "\t-ofolder NAME: output folder, same numbering\n"
               "\t-numfiles X: number of files in the folder\n"
               "\t-numthreads X: number of worker threads, each with its own plan (default 1)\n"
               "\t-gpus X: (extension) spread the worker threads over X CUDA devices (default 1)\n"
               "\t-c2c: (extension) reproduce the reference's C2C branch (what VkResample runs when the upscaled width\n"
               "\t      exceeds its shared-memory limit, e.g. > 6144 on NVIDIA); default is R2C/C2R at every size\n"
               "\t-fast: (extension) approximate divisions / square root in the sharpen (B2R_FLAG_FAST_SHARPEN)\n");
        return 0;
    }
    if (find_flag(argv, argv + argc, "-pngcopy")) {  // diagnostic: decode + re-encode (codec self-test, no GPU)
        char* in = flag_value(argv, argv + argc, "-pngcopy");
        char* out = flag_value(argv, argv + argc, "-o");
        std::vector<unsigned char> rgb; int w, h; std::string err;
        if (!in || !out || !png::load_rgb(in, &rgb, &w, &h, &err)) { printf("pngcopy failed: %s\n", err.c_str()); return 1; }
        return png::write_rgb(out, rgb.data(), w, h) ? 0 : 1;
    }
    if (find_flag(argv, argv + argc, "-devices")) {  // devices_list, VkResample.cpp:239-268
        int n = b2r_device_count();
        for (int i = 0; i < n; ++i) {
            char nm[256];
            b2r_device_name(i, nm, sizeof nm);
            printf("Device id: %d name: %s API:%s\n", i, nm, b2r_version());
        }
        return 0;
    }
    auto need = [&](const char* flag, const char* fmt, void* dst) -> int {
        if (!find_flag(argv, argv + argc, flag)) return 0;
        char* v = flag_value(argv, argv + argc, flag);
        if (!v || sscanf(v, fmt, dst) != 1) { printf("No value is selected with %s flag\n", flag); return 1; }
        return 0;
    };
    if (need("-d", "%u", &cfg.device_id) || need("-n", "%u", &cfg.num_iter) || need("-p", "%u", &cfg.precision) ||
        need("-s", "%f", &cfg.sharpen) || need("-u", "%f", &cfg.upscale) || need("-gpus", "%u", &cfg.gpus))
        return 1;
    cfg.c2c = find_flag(argv, argv + argc, "-c2c") ? 1u : 0u;
    cfg.fast = find_flag(argv, argv + argc, "-fast") ? 1u : 0u;
    if (find_flag(argv, argv + argc, "-ifolder")) {  // batch mode, VkResample.cpp:1893-1957
        cfg.upload_files = 1;
        cfg.ifolder = flag_value(argv, argv + argc, "-ifolder");
        cfg.ofolder = flag_value(argv, argv + argc, "-ofolder");
        if (!cfg.ifolder) { printf("No input folder is selected with -ifolder flag\n"); return 1; }
        if (!cfg.ofolder) { printf("No output folder is selected with -ofolder flag\n"); return 1; }
        if (need("-numthreads", "%u", &cfg.num_threads) || need("-numfiles", "%u", &cfg.num_files)) return 1;
        if (!find_flag(argv, argv + argc, "-numfiles")) { printf("No number of files is selected with -numfiles flag\n"); return 1; }
    } else {
        cfg.input = flag_value(argv, argv + argc, "-i");
        if (!find_flag(argv, argv + argc, "-i") || !cfg.input) { printf("No input file is selected with -i flag\n"); return 1; }
        cfg.output = find_flag(argv, argv + argc, "-o") ? flag_value(argv, argv + argc, "-o") : nullptr;
        cfg.num_threads = 1;
    }
    if (cfg.num_threads < 1) cfg.num_threads = 1;
    if (cfg.gpus < 1) cfg.gpus = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    std::vector<int> rcs(cfg.num_threads, 0);
    for (uint32_t t = 0; t < cfg.num_threads; ++t) {  // VkResample.cpp:1959-1969
        Config c = cfg;
        c.thread_id = t;
        threads.emplace_back([c, &rcs, t] { rcs[t] = launch_resample(c); });
    }
    for (auto& th : threads) th.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Total time: %0.3f s\n", secs);   // VkResample.cpp:1973
    for (int rc : rcs) if (rc) return rc > 0 ? rc : 1;
    return 0;
}
