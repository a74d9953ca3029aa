// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the reference's own lzb_vio::ORBextractor, which is
// compiled from /root/reference/src/ORBextractor.cpp (unmodified, where it lies) against oracle/ref_shim/ by `make -C oracle ref`.
// Output: oracle/_ref/liborb_ref.so (git-ignored).  Used by tests/ to pin the oracle restatement and the CUDA path against
// the reference's compiled code, and by bench.py's CPU-baseline leg.
#include <cstdint>
#include <cstring>
#include <vector>

#include <sys/mman.h>

#include <cstdlib>
#include <new>

#include "lzb_vio/ORBextractor.h"

// ---- allocator switch -------------------------------------------------------------------------------------------------
// DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cpp:636), so nodes with equal key counts are ordered by
// their HEAP ADDRESS: which of them is split last before the `break` depends on the allocator and on the heap's history.
// Mode 0 runs the reference on the system malloc (what a user gets; not reproducible across allocators).  Mode 1 serves every
// allocation of this library from a bump arena that never reuses memory, so addresses grow with creation order -- the
// "monotone allocator" under which the reference's own code realises the canonical tie-break of oracle_orb.cpp and K3.
// (The library is linked with -Bsymbolic, so these replacements are used by this library only.)
namespace {
int g_mode = 0;
char* g_arena = nullptr;
size_t g_off = 0;
const size_t ARENA_BYTES = (size_t)16 << 30;  // virtual; pages are touched on demand
}  // namespace

void* operator new(size_t n) {
    if (g_mode == 1) {
        if (!g_arena) {
            void* p = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) throw std::bad_alloc();
            g_arena = static_cast<char*>(p);
        }
        const size_t a = (g_off + 15) & ~(size_t)15;
        if (a + n > ARENA_BYTES) throw std::bad_alloc();
        g_off = a + n;
        return g_arena + a;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
    if (g_arena && static_cast<char*>(p) >= g_arena && static_cast<char*>(p) < g_arena + ARENA_BYTES) return;  // arena memory is never reused
    std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

extern "C" {

// 0: system malloc, 1: monotone (bump) arena; returns the bytes handed out by the arena so far
size_t ref_set_allocator(int mode) {
    g_mode = mode;
    return g_off;
}

void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new lzb_vio::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}

void ref_orb_destroy(void* h) { delete static_cast<lzb_vio::ORBextractor*>(h); }

// ORBextractor::operator()(image, Mat(), keypoints, descriptors): returns the number of keypoints (or -needed if cap is short);
// kps: 28-byte cv::KeyPoint records, desc: n x 32
int ref_orb_extract(void* h, const uint8_t* img, int w, int h_, size_t stride, void* kps, uint8_t* desc, int cap) {
    lzb_vio::ORBextractor& e = *static_cast<lzb_vio::ORBextractor*>(h);
    cv::Mat image(h_, w, cv::CV_8UC1, const_cast<uint8_t*>(img), stride), mask, d;
    std::vector<cv::KeyPoint> k;
    e(image, mask, k, d);
    const int n = (int)k.size();
    if (n > cap) return -n;
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    if (n) {
        std::memcpy(kps, k.data(), sizeof(cv::KeyPoint) * (size_t)n);
        for (int i = 0; i < n; ++i) std::memcpy(desc + (size_t)32 * i, d.ptr(i), 32);
    }
    return n;
}

// mvImagePyramid[level] (ORBextractor.h:75) of the last extraction, without its border
int ref_orb_level(void* h, int level, uint8_t* dst, size_t dst_stride, int* w, int* h_) {
    lzb_vio::ORBextractor& e = *static_cast<lzb_vio::ORBextractor*>(h);
    if (level < 0 || level >= (int)e.mvImagePyramid.size()) return -1;
    const cv::Mat& m = e.mvImagePyramid[level];
    *w = m.cols;
    *h_ = m.rows;
    if (dst)
        for (int r = 0; r < m.rows; ++r) std::memcpy(dst + (size_t)r * dst_stride, m.ptr(r), (size_t)m.cols);
    return 0;
}

int ref_orb_levels(void* h) { return static_cast<lzb_vio::ORBextractor*>(h)->GetLevels(); }

void ref_orb_scale_factors(void* h, float* out) {
    const std::vector<float> s = static_cast<lzb_vio::ORBextractor*>(h)->GetScaleFactors();
    std::memcpy(out, s.data(), sizeof(float) * s.size());
}
}
