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
//
// Batch mode (-ifolder) has two engines:
//   * -sync: the reference's loop, one frame at a time per worker thread (decode -> upload -> execute ->
//     download -> encode), the worker threads sharing nothing but the GPU (VkResample.cpp:1627-1754);
//   * default: a PIPELINE per GPU -- -numthreads codec workers decode PNGs straight into a ring of pinned
//     host slots and encode finished slots, one submitting thread feeds b2r_enqueue_host_u8 over several
//     lanes (H2D copy, u8->planar, the 4 frame kernels, planar->u8, D2H copy of different frames overlap),
//     completion through b2r_wait_ticket.  Same files, same bytes as -sync (tests/test_cli.py), in the
//     reference's frame -> worker striding across GPUs (frame f, 1-based, -> GPU (f-1) mod gpus).
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
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
// `dst` (optional): caller-provided buffer of dst_cap bytes that receives the pixels instead of *rgb (the
// pipelined batch mode decodes straight into pinned memory); the image must then be exactly want_w x want_h.
static bool load_rgb_impl(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err,
                          unsigned char* dst, size_t dst_cap) {
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
            if (len != 13) { *err = "bad IHDR chunk length"; return false; }
            const uint32_t uw = be32(data), uh = be32(data + 4);
            if (uw == 0 || uh == 0 || uw >= 65536u || uh >= 65536u) { *err = "image dimensions out of range (1..65535)"; return false; }
            width = (int)uw; height = (int)uh;
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0) { *err = "unknown PNG compression / filter method"; return false; }
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
    if (interlace) { *err = "interlaced (Adam7) PNG is not supported"; return false; }
    if (ctype == 3 && plte.empty()) { *err = "palette image without PLTE chunk"; return false; }
    if (idat.empty()) { *err = "no IDAT chunk"; return false; }
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
    unsigned char* pix = dst;
    if (pix) {
        if ((size_t)width * height * 3 != dst_cap) { *err = "image has a different size than the first one"; return false; }
    } else {
        rgb->assign((size_t)width * height * 3, 0);
        pix = rgb->data();
    }
    for (int y = 0; y < height; ++y) {
        const unsigned char* row = &raw[(stride + 1) * y];
        const int ft = row[0];
        const unsigned char* in = row + 1;
        unsigned char* c = cur.data();
        const unsigned char* pr = prev.data();
        switch (ft) {   // one tight loop per filter type (the first bpp bytes have no left neighbour)
            case 0: memcpy(c, in, stride); break;
            case 1:
                for (size_t i = 0; i < bpp && i < stride; ++i) c[i] = in[i];
                for (size_t i = bpp; i < stride; ++i) c[i] = (unsigned char)(in[i] + c[i - bpp]);
                break;
            case 2:
                for (size_t i = 0; i < stride; ++i) c[i] = (unsigned char)(in[i] + pr[i]);
                break;
            case 3:
                for (size_t i = 0; i < bpp && i < stride; ++i) c[i] = (unsigned char)(in[i] + (pr[i] >> 1));
                for (size_t i = bpp; i < stride; ++i) c[i] = (unsigned char)(in[i] + ((c[i - bpp] + pr[i]) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < bpp && i < stride; ++i) c[i] = (unsigned char)(in[i] + pr[i]);
                for (size_t i = bpp; i < stride; ++i) c[i] = (unsigned char)(in[i] + paeth(c[i - bpp], pr[i], pr[i - bpp]));
                break;
            default: *err = "bad PNG filter"; return false;
        }
        unsigned char* out = pix + (size_t)y * width * 3;
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
static bool load_rgb(const char* path, std::vector<unsigned char>* rgb, int* w, int* h, std::string* err,
                     unsigned char* dst = nullptr, size_t dst_cap = 0) {
    try {
        return load_rgb_impl(path, rgb, w, h, err, dst, dst_cap);
    } catch (const std::bad_alloc&) {
        *err = "out of memory while decoding";
        return false;
    }
}

// width / height from the IHDR chunk alone (the pipelined batch mode sizes its pinned ring before decoding anything)
static bool read_size(const char* path, int* w, int* h, std::string* err) {
    FILE* f = fopen(path, "rb");
    if (!f) { *err = "cannot open file"; return false; }
    unsigned char b[33];
    const size_t n = fread(b, 1, sizeof b, f);
    fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 33 || memcmp(b, sig, 8) || be32(b + 8) != 13 || memcmp(b + 12, "IHDR", 4)) { *err = "not a PNG file"; return false; }
    const uint32_t uw = be32(b + 16), uh = be32(b + 20);
    if (uw == 0 || uh == 0 || uw >= 65536u || uh >= 65536u) { *err = "image dimensions out of range (1..65535)"; return false; }
    *w = (int)uw; *h = (int)uh;
    return true;
}

// 8-bit RGB encoder with per-row adaptive filtering (minimum sum of absolute differences)
static int g_zlevel = 6;   // -pnglevel
static bool write_rgb(const char* path, const unsigned char* rgb, int w, int h) {
    const size_t stride = (size_t)w * 3;
    std::vector<unsigned char> raw((stride + 1) * h), cand(stride), best(stride);
    std::vector<unsigned char> zero(stride, 0);
    // candidate rows in tight, branch-free loops (auto-vectorised); cost = sum |signed residual| (the usual heuristic)
    auto cost_of = [&](const unsigned char* r) { long c = 0; for (size_t i = 0; i < stride; ++i) c += std::abs((int)(signed char)r[i]); return c; };
    for (int y = 0; y < h; ++y) {
        const unsigned char* cur = rgb + stride * y;
        const unsigned char* prev = y ? rgb + stride * (y - 1) : zero.data();
        unsigned char* dst = &raw[(stride + 1) * y];
        if (g_zlevel == 0) { dst[0] = 0; memcpy(dst + 1, cur, stride); continue; }   // stored: filtering buys nothing
        long best_cost = -1; int best_ft = 0;
        for (int ft = 0; ft < 5; ++ft) {
            unsigned char* o = cand.data();
            switch (ft) {
                case 0: memcpy(o, cur, stride); break;
                case 1:
                    o[0] = cur[0]; o[1] = cur[1]; o[2] = cur[2];
                    for (size_t i = 3; i < stride; ++i) o[i] = (unsigned char)(cur[i] - cur[i - 3]);
                    break;
                case 2:
                    for (size_t i = 0; i < stride; ++i) o[i] = (unsigned char)(cur[i] - prev[i]);
                    break;
                case 3:
                    for (size_t i = 0; i < 3; ++i) o[i] = (unsigned char)(cur[i] - (prev[i] >> 1));
                    for (size_t i = 3; i < stride; ++i) o[i] = (unsigned char)(cur[i] - ((cur[i - 3] + prev[i]) >> 1));
                    break;
                default:
                    for (size_t i = 0; i < 3; ++i) o[i] = (unsigned char)(cur[i] - prev[i]);
                    for (size_t i = 3; i < stride; ++i) {
                        const int a = cur[i - 3], b = prev[i], c = prev[i - 3];
                        const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
                        const int pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        o[i] = (unsigned char)(cur[i] - pred);
                    }
                    break;
            }
            const long cost = cost_of(o);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_ft = ft; best.swap(cand); }
        }
        dst[0] = (unsigned char)best_ft;
        memcpy(dst + 1, best.data(), stride);
    }
    uLongf zlen = compressBound(raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), raw.size(), g_zlevel) != Z_OK) return false;
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
    uint32_t num_files = 1, gpus = 1, c2c = 0, fast = 0, exact = 0, sync = 0, lanes = 3;
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

static uint32_t plan_flags(const Config& cfg) {
    return (cfg.c2c ? B2R_FLAG_C2C_PARITY : B2R_FLAG_NONE) | (cfg.fast ? B2R_FLAG_FAST_SHARPEN : B2R_FLAG_NONE) |
           (cfg.exact ? B2R_FLAG_EXACT_SHARPEN : B2R_FLAG_NONE);
}

// one-line notices about behaviour a VkResample user would not expect silently (printed once per process)
static void plan_notices(const Config& cfg, const b2r_plan_info& info) {
    static std::atomic<bool> said{false};
    if (said.exchange(true)) return;
    if (!cfg.c2c && info.up_w > 6144)
        printf("Note: VkResample itself switches to its C2C branch for upscaled widths > 6144 on NVIDIA Vulkan "
               "(VkResample.cpp:1424); this run keeps R2C/C2R semantics -- pass -c2c to reproduce that branch.\n");
    if (info.static_kernels != 7u)
        printf("Warning: no statically scheduled kernels for this size (%s); running the any-size kernels at about half speed.\n",
               info.jit_note[0] ? info.jit_note : "plan-time JIT disabled");
}

// launchResample, VkResample.cpp:1280-1780 (the reference's synchronous per-thread frame loop; batch mode: -sync)
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
    if (!png::load_rgb(name, &rgb, &w, &h, &err)) { printf("Image not found (%s: %s)\n", name, err.c_str()); return 5; /* VK_INCOMPLETE */ }

    b2r_plan* plan = nullptr;
    int rc = b2r_plan_create(&plan, device, (uint32_t)w, (uint32_t)h, cfg.upscale, cfg.precision, cfg.sharpen, plan_flags(cfg));
    if (rc) { printf("Plan creation failed, error code: %d (%s)\n", rc, b2r_last_error()); return rc; }
    b2r_plan_info info;
    b2r_plan_get_info(plan, &info);
    plan_notices(cfg, info);
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
            if (!png::load_rgb(name, &rgb, &w2, &h2, &err)) { printf("Image not found (%s: %s)\n", name, err.c_str()); return 5; }
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

// ------------------------------------------------------------------------------- pipelined batch mode
// One Pipeline per GPU.  Frame k of this GPU (global file number first + k*stride, 1-based) lives in ring slot
// k % R from "decode claimed" to "encoded and written":
//     FREE -> DECODING -> DECODED -> SUBMITTED -> (encode claimed) -> FREE
// Codec workers take the oldest unclaimed encode job if there is one (it frees a slot), else the next decode
// job whose slot is free; the submitting thread hands DECODED slots to b2r_enqueue_host_u8 in frame order and
// records the ticket; an encoder waits for its ticket (b2r_wait_ticket, thread-safe) before reading the slot.
struct Pipeline {
    Config cfg;
    int gpu_index = 0, device = 0, w = 0, h = 0;
    uint32_t first = 1, stride = 1, total = 0;       // this GPU's files: first, first+stride, ...
    b2r_plan* plan = nullptr;
    b2r_plan_info info{};
    enum State { FREE, DECODING, DECODED, SUBMITTED };
    struct Slot { unsigned char* in = nullptr; unsigned char* out = nullptr; State state = FREE; uint64_t ticket = 0; };
    std::vector<Slot> slots;
    std::mutex mu;
    std::condition_variable cv;
    uint32_t next_decode = 0, next_submit = 0, next_encode = 0, written = 0;
    int error = 0;
    double t_decode = 0, t_encode = 0, t_wait = 0;    // summed over workers (seconds)
    std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    double at_ring = 0, at_plan = 0, at_first = 0, at_last = 0;   // seconds since t_start: ring allocated, plan ready, first / last file written
    double since_start() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); }

    uint32_t file_of(uint32_t k) const { return first + k * stride; }

    void fail_with(int rc) { std::lock_guard<std::mutex> g(mu); if (!error) error = rc ? rc : -1; cv.notify_all(); }

    void submit_loop() {
        for (uint32_t k = 0; k < total; ++k) {
            Slot& s = slots[k % slots.size()];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return error || s.state == DECODED; });
                if (error) return;
            }
            int rc = b2r_enqueue_host_u8(plan, s.in, s.out);
            if (rc) { printf("enqueue failed: %s\n", b2r_last_error()); fail_with(rc); return; }
            std::lock_guard<std::mutex> g(mu);
            s.ticket = b2r_plan_last_ticket(plan);
            s.state = SUBMITTED;
            next_submit = k + 1;
            cv.notify_all();
        }
    }

    void worker_loop() {
        using clk = std::chrono::steady_clock;
        char name[512];
        std::string err;
        for (;;) {
            uint32_t k = 0; bool encode = false;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    if (error || written == total) return true;
                    if (next_encode < next_submit) return true;
                    return next_decode < total && slots[next_decode % slots.size()].state == FREE;
                });
                if (error || written == total) return;
                if (next_encode < next_submit) { k = next_encode++; encode = true; }
                else { k = next_decode++; slots[k % slots.size()].state = DECODING; }
            }
            Slot& s = slots[k % slots.size()];
            if (encode) {
                auto t0 = clk::now();
                int rc = b2r_wait_ticket(plan, s.ticket);
                auto t1 = clk::now();
                if (rc) { printf("wait failed: %s\n", b2r_last_error()); fail_with(rc); return; }
                snprintf(name, sizeof name, "%s/%06u.png", cfg.ofolder, file_of(k));
                if (!png::write_rgb(name, s.out, (int)info.up_w, (int)info.up_h)) printf("cannot write %s\n", name);
                auto t2 = clk::now();
                std::lock_guard<std::mutex> g(mu);
                t_wait += std::chrono::duration<double>(t1 - t0).count();
                t_encode += std::chrono::duration<double>(t2 - t1).count();
                s.state = FREE;
                ++written;
                at_last = since_start();
                if (written == 1) at_first = at_last;
                cv.notify_all();
            } else {
                auto t0 = clk::now();
                if (!s.in) {   // first use of this ring slot
                    s.in = (unsigned char*)b2r_host_alloc((size_t)3 * w * h);
                    s.out = (unsigned char*)b2r_host_alloc((size_t)3 * info.up_w * info.up_h);
                    if (!s.in || !s.out) { printf("pinned allocation failed: %s\n", b2r_last_error()); fail_with(-4); return; }
                }
                snprintf(name, sizeof name, "%s/%06u.png", cfg.ifolder, file_of(k));
                int w2 = 0, h2 = 0;
                if (!png::load_rgb(name, nullptr, &w2, &h2, &err, s.in, (size_t)w * h * 3)) {
                    printf("Image not found (%s: %s)\n", name, err.c_str());
                    fail_with(5);
                    return;
                }
                std::lock_guard<std::mutex> g(mu);
                t_decode += std::chrono::duration<double>(clk::now() - t0).count();
                s.state = DECODED;
                cv.notify_all();
            }
        }
    }

    // plan creation runs on the submitting thread while the codec workers already decode into the ring
    int create_plan() {
        int rc = b2r_plan_create(&plan, device, (uint32_t)w, (uint32_t)h, cfg.upscale, cfg.precision, cfg.sharpen, plan_flags(cfg));
        if (rc) { printf("Plan creation failed, error code: %d (%s)\n", rc, b2r_last_error()); return rc; }
        const uint32_t lanes = std::max(1u, std::min(cfg.lanes, 8u));
        if ((rc = b2r_plan_set_lanes(plan, lanes))) { printf("set_lanes failed: %s\n", b2r_last_error()); return rc; }
        b2r_plan_info pi;
        b2r_plan_get_info(plan, &pi);
        if (pi.up_w != info.up_w || pi.up_h != info.up_h) { printf("internal error: output size mismatch\n"); return -1; }
        plan_notices(cfg, pi);
        if (gpu_index == 0)
            printf("VRAM per GPU: %d MB (%u lanes), %u codec threads + 1 submit thread per GPU\n", (int)(pi.device_bytes >> 20), lanes, cfg.num_threads);
        std::lock_guard<std::mutex> g(mu);
        info = pi;
        at_plan = since_start();
        return 0;
    }

    int run() {
        const int ndev = b2r_device_count();
        if (ndev < 1) { printf("No CUDA device found: %s\n", b2r_last_error()); return -1; }
        device = (int)((cfg.device_id + (uint32_t)gpu_index) % (uint32_t)ndev);
        char name[512];
        snprintf(name, sizeof name, "%s/%06u.png", cfg.ifolder, file_of(0));
        std::string err;
        if (!png::read_size(name, &w, &h, &err)) { printf("Image not found (%s: %s)\n", name, err.c_str()); return 5; }
        // output size as the reference computes it (VkResample.cpp:1417-1418); checked against the plan's once it exists
        info.up_w = (uint32_t)(cfg.upscale * (float)w); info.up_h = (uint32_t)(cfg.upscale * (float)h);
        const uint32_t lanes = std::max(1u, std::min(cfg.lanes, 8u));
        const size_t n_slots = std::min<size_t>(total, (size_t)cfg.num_threads + 2 * lanes + 2);
        slots.resize(n_slots);   // pinned buffers are allocated by the codec worker that first decodes into a slot (in parallel)
        at_ring = since_start();
        std::vector<std::thread> th;
        th.emplace_back([this] { int rc = create_plan(); if (rc) fail_with(rc); else submit_loop(); });
        for (uint32_t t = 0; t < cfg.num_threads; ++t) th.emplace_back([this] { worker_loop(); });
        for (auto& t : th) t.join();
        if (plan) b2r_synchronize(plan);
        {   // release the ring in parallel (page unpinning is the slow part)
            std::vector<std::thread> fr;
            for (size_t t0 = 0; t0 < slots.size(); t0 += 4)
                fr.emplace_back([this, t0] { for (size_t i = t0; i < std::min(t0 + 4, slots.size()); ++i) { b2r_host_free(slots[i].in); b2r_host_free(slots[i].out); } });
            for (auto& t : fr) t.join();
        }
        char dev_name[256] = "";
        b2r_device_name(device, dev_name, sizeof dev_name);
        printf("GPU %d finished: %u frames, codec thread-seconds: decode %.2f, encode %.2f, waiting for the GPU %.2f. Device name: %s API:%s\n",
               gpu_index, written, t_decode, t_encode, t_wait, dev_name, b2r_version());
        printf("GPU %d timeline: pinned ring ready at %.2f s, plan ready at %.2f s, first file written at %.2f s, last at %.2f s",
               gpu_index, at_ring, at_plan, at_first, at_last);
        if (written > 1 && at_last > at_first) printf(" -> steady state %.1f frames/s", (written - 1) / (at_last - at_first));
        printf("\n");
        if (plan) b2r_plan_destroy(plan);
        return error;
    }
};

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
               "\t-exact: (extension) bit-exact sharpen arithmetic (B2R_FLAG_EXACT_SHARPEN; default: tolerance-bound kernels)\n"
               "\t-fast: (round-1 extension, now the default behaviour; kept for compatibility)\n"
               "\t-sync: (extension) batch mode as the reference runs it: every worker thread decodes, uploads, executes,\n"
               "\t       downloads and encodes one frame at a time.  Default batch mode is a pipeline per GPU: -numthreads\n"
               "\t       codec threads around a pinned ring, frames in flight on -lanes X streams (default 3)\n"
               "\t-pnglevel X: (extension) zlib level of the PNG encoder, 0..9 (default 6)\n");
        return 0;
    }
    if (char* lv = flag_value(argv, argv + argc, "-pnglevel")) { int z = atoi(lv); if (z >= 0 && z <= 9) png::g_zlevel = z; }
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
    cfg.exact = find_flag(argv, argv + argc, "-exact") ? 1u : 0u;
    cfg.sync = find_flag(argv, argv + argc, "-sync") ? 1u : 0u;
    if (need("-lanes", "%u", &cfg.lanes) || need("-pnglevel", "%d", &png::g_zlevel)) return 1;
    if (png::g_zlevel < 0 || png::g_zlevel > 9) png::g_zlevel = 6;
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
    if (cfg.upload_files && !cfg.sync) {   // pipelined batch mode: one Pipeline per GPU, frame f -> GPU (f-1) mod gpus
        printf("VkResample - FFT based upscaling\n");
        const uint32_t g_used = std::min(cfg.gpus, std::max(1u, cfg.num_files));
        std::vector<Pipeline> pipes(g_used);
        std::vector<std::thread> th;
        std::vector<int> prc(g_used, 0);
        for (uint32_t g = 0; g < g_used; ++g) {
            Pipeline& p = pipes[g];
            p.cfg = cfg; p.gpu_index = (int)g; p.first = g + 1; p.stride = g_used;
            p.total = (cfg.num_files > g) ? (cfg.num_files - g - 1) / g_used + 1 : 0;   // VkResample.cpp:1622-1626 with numThreads = gpus
            th.emplace_back([&p, &prc, g] { prc[g] = p.total ? p.run() : 0; });
        }
        for (auto& t : th) t.join();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint32_t frames = 0;
        for (auto& p : pipes) frames += p.written;
        printf("Total time: %0.3f s\n", secs);
        printf("Pipelined batch: %u frames on %u GPU(s), %.1f frames/s PNG in -> PNG out (plan creation included)\n", frames, g_used, frames / secs);
        for (int rc : prc) if (rc) return rc > 0 ? rc : 1;
        return 0;
    }
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
