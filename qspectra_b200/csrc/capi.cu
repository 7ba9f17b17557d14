// Library-wide pieces of the C ABI: error reporting, launch accounting, device
// query and the closed-form ADO enumeration (integer index maps of the HEOM
// hierarchy, bit-exact with the reference's ADO_mappings / multichoose,
// dynamics/heom.py:92-174).
#include "common.cuh"
#include "ado.h"
#include <string.h>

static thread_local char g_error[1024] = "";
std::atomic<uint64_t> qsx_launch_counter{0};

void qsx_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char *qsx_last_error(void) { return g_error; }
extern "C" int qsx_version(void) { return 100; }
extern "C" uint64_t qsx_kernel_launches(void) { return qsx_launch_counter.load(); }

extern "C" int qsx_device_info(int32_t *sm_count, int64_t *l2_bytes, int32_t *smem_per_block) {
    int dev = 0;
    QSX_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    QSX_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    if (smem_per_block) *smem_per_block = (int32_t)prop.sharedMemPerBlockOptin;
    return QSX_OK;
}

// --------------------------------------------------------------- ADO tables
AdoTables::AdoTables(int bins_, int level_cutoff_) : bins(bins_), level_cutoff(level_cutoff_) {
    // Pascal triangle up to (bins + level_cutoff)
    int top = bins + level_cutoff + 2;
    binom.assign((size_t)top * top, 0);
    for (int a = 0; a < top; ++a) {
        binom[(size_t)a * top] = 1;
        for (int b = 1; b <= a; ++b)
            binom[(size_t)a * top + b] = binom[(size_t)(a - 1) * top + b - 1] +
                                         (b <= a - 1 ? binom[(size_t)(a - 1) * top + b] : 0);
    }
    btop = top;
    level_offset.assign(level_cutoff + 1, 0);
    for (int c = 0; c < level_cutoff; ++c)
        level_offset[c + 1] = level_offset[c] + (bins > 0 ? C(bins + c - 1, c) : (c == 0));
    n_ado = level_offset[level_cutoff];
}

int64_t AdoTables::rank(const int *v) const {
    int c = 0;
    for (int i = 0; i < bins; ++i) c += v[i];
    if (c < 0 || c >= level_cutoff) return -1;
    int64_t r = level_offset[c];
    int rem = c;
    for (int i = 0; i < bins; ++i) {
        int tail = bins - i - 1;           // bins after position i
        if (tail == 0) break;
        for (int x = 0; x < v[i]; ++x) r += C(rem - x + tail - 1, tail - 1);
        rem -= v[i];
    }
    return r;
}

void AdoTables::enumerate() {
    index.assign((size_t)n_ado * bins, 0);
    std::vector<int> v(bins, 0);
    int64_t row = 0;
    for (int c = 0; c < level_cutoff; ++c) {
        // first composition of c in lexicographic order: (0, ..., 0, c)
        std::fill(v.begin(), v.end(), 0);
        if (bins > 0) v[bins - 1] = c;
        while (true) {
            for (int i = 0; i < bins; ++i) index[(size_t)row * bins + i] = (uint8_t)v[i];
            ++row;
            // successor in lexicographic order
            if (bins < 2) break;
            if (v[bins - 1] > 0) {
                v[bins - 2] += 1;
                v[bins - 1] -= 1;
                continue;
            }
            int p = bins - 2;
            while (p >= 0 && v[p] == 0) --p;
            if (p <= 0) break;          // all mass in position 0 (or c == 0): last one
            int mass = v[p] - 1;
            v[p] = 0;
            v[p - 1] += 1;
            v[bins - 1] = mass;
        }
    }
    up.assign((size_t)n_ado * bins, -1);
    down.assign((size_t)n_ado * bins, -1);
    for (int64_t n = 0; n < n_ado; ++n) {
        int c = 0;
        for (int i = 0; i < bins; ++i) { v[i] = index[(size_t)n * bins + i]; c += v[i]; }
        for (int b = 0; b < bins; ++b) {
            if (c + 1 < level_cutoff) {
                v[b] += 1;
                up[(size_t)n * bins + b] = (int32_t)rank(v.data());
                v[b] -= 1;
            }
            if (v[b] > 0) {
                v[b] -= 1;
                down[(size_t)n * bins + b] = (int32_t)rank(v.data());
                v[b] += 1;
            }
        }
    }
}

extern "C" int64_t qsx_ado_count(int32_t bins, int32_t level_cutoff) {
    if (bins < 0 || level_cutoff < 0) return -1;
    AdoTables t(bins, level_cutoff);
    return t.n_ado;
}

extern "C" int qsx_ado_enumerate(int32_t bins, int32_t level_cutoff, int64_t *ado_index,
                                 int32_t *up, int32_t *down) {
    QSX_REQUIRE(bins > 0 && level_cutoff > 0, "qsx_ado_enumerate: bad arguments");
    AdoTables t(bins, level_cutoff);
    QSX_REQUIRE(t.n_ado < (int64_t)1 << 31, "hierarchy too large for 32-bit neighbour tables");
    t.enumerate();
    size_t total = (size_t)t.n_ado * bins;
    if (ado_index) for (size_t i = 0; i < total; ++i) ado_index[i] = t.index[i];
    if (up) memcpy(up, t.up.data(), total * sizeof(int32_t));
    if (down) memcpy(down, t.down.data(), total * sizeof(int32_t));
    return QSX_OK;
}
