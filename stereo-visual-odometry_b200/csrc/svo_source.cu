// svo_source.cu -- frame source of the throughput mode (host code).
//
// Replaces System::NextFrame_kitti (System.cpp:75-104): cv::imread("<dataset>/image_0/%06d.png", IMREAD_GRAYSCALE) for the left and
// image_1 for the right image of frame current_image_index_, nullptr (end of the run) at the first missing file.  The reference
// decodes one pair per Step() on the tracking thread; here a pool of decoder threads fills a ring of PINNED batches ahead of the
// consumer, so that svo_frontend_upload() can take whole batches straight from the ring at PCIe rate.
//
// Formats: PNG (8-bit greyscale, non-interlaced: what the KITTI odometry grey sequences are; inflate by zlib, the five PNG row
// filters undone here) and binary PGM (P5, maxval 255), looked up as %06d.png then %06d.pgm.  Anything else is an error, never
// a silent conversion.  No pixel arithmetic of the hot path happens here.
#include <cuda_runtime.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "svo_cuda.h"

namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == buf.size();
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// 8-bit greyscale, non-interlaced PNG -> w x h bytes at dst (row stride w).  Returns an empty string on success.
std::string decode_png(const std::vector<uint8_t>& file, int w, int h, uint8_t* dst, std::vector<uint8_t>& scratch) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) return "not a PNG file";
    size_t pos = 8;
    bool have_hdr = false;
    std::vector<uint8_t>& idat = scratch;
    idat.clear();
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return "truncated PNG chunk";
        const uint8_t* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return "bad IHDR";
            const uint32_t pw = be32(data), ph = be32(data + 4);
            if ((int)pw != w || (int)ph != h) return "PNG size " + std::to_string(pw) + "x" + std::to_string(ph) + " differs from the configured size";
            if (data[8] != 8 || data[9] != 0) return "only 8-bit greyscale PNG is supported (bit depth " + std::to_string(data[8]) + ", colour type " + std::to_string(data[9]) + ")";
            if (data[12] != 0) return "interlaced PNG is not supported";
            have_hdr = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_hdr || idat.empty()) return "PNG without IHDR / IDAT";
    const size_t raw_bytes = (size_t)(w + 1) * h;  // one filter byte per row
    std::vector<uint8_t> raw(raw_bytes);
    uLongf out_len = (uLongf)raw_bytes;
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw_bytes) return "PNG inflate failed";
    // undo the row filters (bytes per pixel = 1)
    const uint8_t* prev = nullptr;
    for (int y = 0; y < h; ++y) {
        const uint8_t* in = &raw[(size_t)y * (w + 1)];
        uint8_t* out = dst + (size_t)y * w;
        const int ft = in[0];
        ++in;
        switch (ft) {
            case 0: memcpy(out, in, (size_t)w); break;
            case 1:
                out[0] = in[0];
                for (int x = 1; x < w; ++x) out[x] = (uint8_t)(in[x] + out[x - 1]);
                break;
            case 2:
                for (int x = 0; x < w; ++x) out[x] = (uint8_t)(in[x] + (prev ? prev[x] : 0));
                break;
            case 3:
                for (int x = 0; x < w; ++x) out[x] = (uint8_t)(in[x] + (((x ? out[x - 1] : 0) + (prev ? prev[x] : 0)) >> 1));
                break;
            case 4:
                for (int x = 0; x < w; ++x) {
                    const int a = x ? out[x - 1] : 0, b = prev ? prev[x] : 0, c = (x && prev) ? prev[x - 1] : 0;
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    out[x] = (uint8_t)(in[x] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
                }
                break;
            default: return "unknown PNG row filter";
        }
        prev = out;
    }
    return "";
}

std::string decode_pgm(const std::vector<uint8_t>& file, int w, int h, uint8_t* dst) {
    // "P5" whitespace width whitespace height whitespace maxval ONE whitespace, then w*h bytes; '#' comments allowed in the header
    size_t pos = 0;
    auto token = [&]() -> std::string {
        for (;;) {
            while (pos < file.size() && isspace(file[pos])) ++pos;
            if (pos < file.size() && file[pos] == '#') {
                while (pos < file.size() && file[pos] != '\n') ++pos;
                continue;
            }
            break;
        }
        std::string t;
        while (pos < file.size() && !isspace(file[pos])) t.push_back((char)file[pos++]);
        return t;
    };
    if (token() != "P5") return "not a binary PGM (P5) file";
    const int pw = atoi(token().c_str()), ph = atoi(token().c_str()), maxv = atoi(token().c_str());
    if (pw != w || ph != h) return "PGM size " + std::to_string(pw) + "x" + std::to_string(ph) + " differs from the configured size";
    if (maxv != 255) return "only 8-bit PGM (maxval 255) is supported";
    ++pos;  // the single whitespace after maxval
    if (pos + (size_t)w * h > file.size()) return "truncated PGM";
    memcpy(dst, &file[pos], (size_t)w * h);
    return "";
}

bool exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

}  // namespace

struct svo_source {
    std::string dir, err;
    int first = 0, w = 0, h = 0, batch = 0, ring = 0, n_total = 0;
    bool pinned = false;
    uint8_t* mem = nullptr;  // ring x 2 sides x batch x h x w
    size_t side_bytes = 0;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_ready, cv_free;
    std::vector<int> remaining;   // per ring slot: decode jobs left of the batch it currently holds (-1: free)
    std::vector<int> slot_batch;  // which batch the slot holds
    int next_job = 0;             // next (frame, side) job to hand to a worker: job = 2 * frame + side
    int consumed_batches = 0;     // batches handed to the consumer so far
    int held_slot[2] = {-1, -1};  // slots lent to the consumer by the last two svo_source_next() calls ([0] = the older one)
    bool stop = false, failed = false;
    std::atomic<long long> decoded_bytes{0};

    std::string path(int side, int idx, const char* ext) const {
        char name[64];
        snprintf(name, sizeof name, "/image_%d/%06d.%s", side, idx, ext);
        return dir + name;
    }
    uint8_t* slot_ptr(int slot, int side) const { return mem + ((size_t)slot * 2 + side) * side_bytes; }

    void work() {
        std::vector<uint8_t> file, scratch;
        for (;;) {
            int job;
            {
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    if (stop || failed) return;
                    if (next_job >= 2 * n_total) return;
                    const int b = (next_job / 2) / batch, slot = b % ring;
                    if (slot_batch[slot] == b) break;                        // the slot already holds this batch: take the job
                    if (remaining[slot] == -1 && b < consumed_batches + ring) {  // free slot, and its turn: claim it for batch b
                        slot_batch[slot] = b;
                        remaining[slot] = 2 * std::min(batch, n_total - b * batch);
                        break;
                    }
                    cv_free.wait(lk);
                }
                job = next_job++;
            }
            const int frame = job / 2, side = job & 1, b = frame / batch, slot = b % ring;
            uint8_t* dst = slot_ptr(slot, side) + (size_t)(frame - b * batch) * w * h;
            std::string e;
            if (read_file(path(side, first + frame, "png"), file)) e = decode_png(file, w, h, dst, scratch);
            else if (read_file(path(side, first + frame, "pgm"), file)) e = decode_pgm(file, w, h, dst);
            else e = "cannot read the file";
            decoded_bytes += (long long)w * h;
            std::unique_lock<std::mutex> lk(mu);
            if (!e.empty()) {
                if (!failed) err = path(side, first + frame, "png|pgm") + ": " + e;
                failed = true;
                cv_ready.notify_all();
                cv_free.notify_all();
                return;
            }
            if (--remaining[slot] == 0) cv_ready.notify_all();
        }
    }
};

extern "C" int svo_source_open(svo_source** out, const char* dir, int first_index, int w, int h, int batch, int ring_batches, int threads) {
    if (!out) return SVO_ERR_INVALID;
    *out = nullptr;
    if (!dir || first_index < 0 || w <= 0 || h <= 0 || batch < 1 || ring_batches < 3 || threads < 1) return SVO_ERR_INVALID;
    svo_source* s = new svo_source();
    s->dir = dir;
    s->first = first_index;
    s->w = w;
    s->h = h;
    s->batch = batch;
    s->ring = ring_batches;
    // the run ends at the first missing pair, like the reference's nullptr frame (System.cpp:88-93)
    while (true) {
        const int idx = first_index + s->n_total;
        const bool l = exists(s->path(0, idx, "png")) || exists(s->path(0, idx, "pgm"));
        const bool r = exists(s->path(1, idx, "png")) || exists(s->path(1, idx, "pgm"));
        if (!l || !r) break;
        ++s->n_total;
    }
    s->side_bytes = (size_t)batch * w * h;
    const size_t total = (size_t)ring_batches * 2 * s->side_bytes;
    // pinned memory so that svo_frontend_upload() copies at full PCIe rate; on a host without a CUDA device (tools, CPU tests of the
    // decoders) ordinary memory holds the ring -- this is staging memory, not a compute path
    void* p = nullptr;
    if (cudaHostAlloc(&p, total, cudaHostAllocDefault) == cudaSuccess) {
        s->pinned = true;
    } else {
        cudaGetLastError();
        if (posix_memalign(&p, 4096, total) != 0) {
            delete s;
            return SVO_ERR_INVALID;
        }
    }
    s->mem = (uint8_t*)p;
    s->remaining.assign(ring_batches, -1);
    s->slot_batch.assign(ring_batches, -1);
    for (int i = 0; i < threads; ++i) s->workers.emplace_back([s] { s->work(); });
    *out = s;
    return SVO_OK;
}

extern "C" int svo_source_frame_count(const svo_source* s) { return s ? s->n_total : 0; }
extern "C" int svo_source_is_pinned(const svo_source* s) { return s && s->pinned ? 1 : 0; }
extern "C" const char* svo_source_error(const svo_source* s) { return s ? s->err.c_str() : "null source"; }

extern "C" int svo_source_next(svo_source* s, const uint8_t** left, const uint8_t** right, int* n_frames) {
    if (!s || !left || !right || !n_frames) return SVO_ERR_INVALID;
    *left = *right = nullptr;
    *n_frames = 0;
    std::unique_lock<std::mutex> lk(s->mu);
    // A lent batch stays valid for TWO calls: its asynchronous host-to-device copy (svo_frontend_upload) may still be reading it when
    // the next batch is asked for; one call later the pipelined caller has downloaded that batch's results, so its copy is complete.
    if (s->held_slot[0] >= 0) {
        s->remaining[s->held_slot[0]] = -1;
        s->slot_batch[s->held_slot[0]] = -1;
        s->cv_free.notify_all();
    }
    s->held_slot[0] = s->held_slot[1];
    s->held_slot[1] = -1;
    const int b = s->consumed_batches;
    if (b * s->batch >= s->n_total) return s->failed ? SVO_ERR_INVALID : SVO_OK;  // end of the sequence: n_frames = 0
    const int slot = b % s->ring;
    s->cv_ready.wait(lk, [&] { return s->failed || (s->slot_batch[slot] == b && s->remaining[slot] == 0); });
    if (s->failed) return SVO_ERR_INVALID;
    *left = s->slot_ptr(slot, 0);
    *right = s->slot_ptr(slot, 1);
    *n_frames = std::min(s->batch, s->n_total - b * s->batch);
    s->held_slot[1] = slot;
    ++s->consumed_batches;
    s->cv_free.notify_all();  // the window of claimable batches moved
    return SVO_OK;
}

extern "C" void svo_source_close(svo_source* s) {
    if (!s) return;
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->stop = true;
        s->cv_free.notify_all();
        s->cv_ready.notify_all();
    }
    for (auto& t : s->workers) t.join();
    if (s->mem) {
        if (s->pinned) cudaFreeHost(s->mem);
        else free(s->mem);
    }
    delete s;
}

// ---- writers (tests / bench: synthetic sequences on disk in the reference's dataset layout) --------------------------------
extern "C" int svo_write_pgm(const char* path, const uint8_t* img, int w, int h, size_t stride) {
    if (!path || !img || w <= 0 || h <= 0 || stride < (size_t)w) return SVO_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return SVO_ERR_INVALID;
    fprintf(f, "P5\n%d %d\n255\n", w, h);
    for (int y = 0; y < h; ++y) fwrite(img + (size_t)y * stride, 1, (size_t)w, f);
    fclose(f);
    return SVO_OK;
}

extern "C" int svo_write_png(const char* path, const uint8_t* img, int w, int h, size_t stride, int level) {
    if (!path || !img || w <= 0 || h <= 0 || stride < (size_t)w) return SVO_ERR_INVALID;
    // filter 1 (Sub) on every row: what typical encoders pick for photographs, and it exercises the decoder's defiltering
    std::vector<uint8_t> raw((size_t)(w + 1) * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* in = img + (size_t)y * stride;
        uint8_t* o = &raw[(size_t)y * (w + 1)];
        const int ft = (y % 5 == 4) ? 2 : 1;  // mostly Sub, every fifth row Up
        o[0] = (uint8_t)ft;
        if (ft == 1) {
            o[1] = in[0];
            for (int x = 1; x < w; ++x) o[1 + x] = (uint8_t)(in[x] - in[x - 1]);
        } else {
            const uint8_t* up = in - stride;
            for (int x = 0; x < w; ++x) o[1 + x] = (uint8_t)(in[x] - up[x]);
        }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), level) != Z_OK) return SVO_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return SVO_ERR_INVALID;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    fwrite(sig, 1, 8, f);
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1],
                          (uint8_t)type[2], (uint8_t)type[3]};
        fwrite(hdr, 1, 8, f);
        if (len) fwrite(data, 1, len, f);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        fwrite(c, 1, 4, f);
    };
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16),
                        (uint8_t)(h >> 8), (uint8_t)h, 8, 0, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    fclose(f);
    return SVO_OK;
}
