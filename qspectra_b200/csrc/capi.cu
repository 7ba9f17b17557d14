// Library-wide pieces of the C ABI: error reporting, launch accounting, device
// query and the closed-form ADO enumeration (integer index maps of the HEOM
// hierarchy, bit-exact with the reference's ADO_mappings / multichoose,
// dynamics/heom.py:92-174).
#include "common.cuh"
#include "ado.h"
#include <string.h>
#include <math.h>
#include <algorithm>
#include <thread>
#include <map>
#include <mutex>
#include <vector>

static thread_local char g_error[1024] = "";
std::atomic<uint64_t> qsx_launch_counter{0};
std::atomic<uint64_t> qsx_h2d_counter{0}, qsx_d2h_counter{0};

void qsx_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------ device scratch pool
namespace {
struct Pool {
    std::mutex mu;
    std::map<void *, size_t> live;                       // ptr -> size class (0: not pooled)
    std::map<std::pair<int, size_t>, std::vector<void *>> idle;   // (device, class) -> blocks
};
Pool &pool() { static Pool p; return p; }
const size_t kPoolMax = (size_t)64 << 20;
}  // namespace

cudaError_t qsx_pool_alloc(void **ptr, size_t bytes) {
    size_t cls = 256;
    while (cls < bytes) cls <<= 1;
    int dev = 0;
    cudaGetDevice(&dev);
    Pool &P = pool();
    if (bytes <= kPoolMax) {
        std::lock_guard<std::mutex> lock(P.mu);
        auto &v = P.idle[std::make_pair(dev, cls)];
        if (!v.empty()) {
            *ptr = v.back();
            v.pop_back();
            P.live[*ptr] = cls;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(ptr, bytes <= kPoolMax ? cls : bytes);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(P.mu);
        P.live[*ptr] = bytes <= kPoolMax ? cls : 0;
    }
    return e;
}

void qsx_pool_free(void *ptr) {
    if (!ptr) return;
    Pool &P = pool();
    size_t cls = 0;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.live.find(ptr);
        if (it != P.live.end()) { cls = it->second; P.live.erase(it); }
        if (cls) {
            int dev = 0;
            cudaGetDevice(&dev);
            P.idle[std::make_pair(dev, cls)].push_back(ptr);
            return;
        }
    }
    cudaFree(ptr);
}

extern "C" const char *qsx_last_error(void) { return g_error; }
extern "C" int qsx_version(void) { return 100; }
extern "C" uint64_t qsx_kernel_launches(void) { return qsx_launch_counter.load(); }
extern "C" void qsx_transfer_bytes(uint64_t *h2d, uint64_t *d2h) {
    if (h2d) *h2d = qsx_h2d_counter.load();
    if (d2h) *d2h = qsx_d2h_counter.load();
}

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

// ------------------------------------------------------------ K6: member sum
// Deterministic two-pass sum: pass 1 writes the partial sum of each member block, pass 2 adds the
// partials of an element in block order (no floating-point atomics: ensemble means are
// bit-reproducible from run to run, like the seed-replayed members themselves).
__global__ void reduce_members_kernel(const cplx *__restrict__ in, int n_members, long long n,
                                      int members_per_block, double scale, cplx *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m0 = blockIdx.y * members_per_block;
    int m1 = min(n_members, m0 + members_per_block);
    double sr = 0.0, si = 0.0;
    for (int m = m0; m < m1; ++m) {
        cplx v = __ldg(&in[(size_t)m * n + i]);
        sr += v.x;
        si += v.y;
    }
    out[(size_t)blockIdx.y * n + i] = cmake(scale * sr, scale * si);
}
__global__ void reduce_partials_kernel(const cplx *__restrict__ part, int n_part, long long n,
                                       cplx *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sr = 0.0, si = 0.0;
    for (int p = 0; p < n_part; ++p) {
        cplx v = part[(size_t)p * n + i];
        sr += v.x;
        si += v.y;
    }
    out[i] = cmake(sr, si);
}

extern "C" int qsx_reduce_members(const void *in_dev, int32_t n_members, int64_t n, double scale,
                                  void *out_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(in_dev && out_dev && n_members > 0 && n > 0, "qsx_reduce_members: bad arguments");
    const int threads = 128;
    long long bx = (n + threads - 1) / threads;
    // enough blocks to fill the machine, at most one per 8 members
    int by = (int)std::max<long long>(1, std::min<long long>((n_members + 7) / 8, (148 * 16 + bx - 1) / bx));
    int per = (n_members + by - 1) / by;
    by = (n_members + per - 1) / per;
    if (by == 1) {
        reduce_members_kernel<<<dim3((unsigned)bx, 1), threads, 0, stream>>>(
            (const cplx *)in_dev, n_members, n, per, scale, (cplx *)out_dev);
        qsx_launch_counter += 1;
    } else {
        DevBuf<cplx> part;
        QSX_CUDA(part.alloc((size_t)by * n));
        reduce_members_kernel<<<dim3((unsigned)bx, (unsigned)by), threads, 0, stream>>>(
            (const cplx *)in_dev, n_members, n, per, scale, part.p);
        reduce_partials_kernel<<<(unsigned)bx, threads, 0, stream>>>(part.p, by, n, (cplx *)out_dev);
        qsx_launch_counter += 2;
    }
    QSX_CUDA(cudaGetLastError());
    return QSX_OK;
}

// ------------------------------------------------- seeded disorder streams (host)
// Bit-exact replay of numpy's legacy generator for ensemble member n:
//   rng = np.random.RandomState(list(seed) + [n]); rng.randn(n_gauss); rng.rand(n_uniform)
// (reference hamiltonian.py:458-461, 566-573).  MT19937 init_by_array + the
// legacy polar Box-Muller with its one-value cache, as in numpy's
// _legacy/mt19937 sources; ~100x faster than constructing RandomState objects.
namespace {
struct MT19937 {
    uint32_t mt[624];
    int pos;
    __host__ __device__ void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253U * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
    __host__ __device__ void init_by_array(const uint32_t *key, int len) {
        init_genrand(19650218U);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000U;
        pos = 624;
    }
    __host__ __device__ void twist() {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000U) | (mt[(k + 1) % 624] & 0x7fffffffU);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
        }
        pos = 0;
    }
    __host__ __device__ uint32_t next32() {
        if (pos >= 624) twist();
        uint32_t y = mt[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    }
    __host__ __device__ double next_double() {
        uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};
}  // namespace

extern "C" int qsx_sample_streams(const uint32_t *seed_prefix, int32_t n_prefix, int64_t member0,
                                  int32_t n_members, int32_t n_gauss, int32_t n_uniform,
                                  double *gauss_out, double *uniform_out) {
    QSX_REQUIRE(n_prefix >= 0 && n_prefix < 15 && n_members >= 0 && n_gauss >= 0 && n_uniform >= 0,
                "qsx_sample_streams: bad arguments");
    QSX_REQUIRE(member0 >= 0 && member0 + n_members <= (int64_t)0xffffffffLL, "member index out of range");
    // members are independent streams: spread them over the host cores
    unsigned hw = std::thread::hardware_concurrency();
    int n_threads = (int)std::max(1u, std::min(hw ? hw : 1u, (unsigned)((n_members + 255) / 256)));
    auto work = [&](int lo, int hi) {
        uint32_t key[16];
        for (int i = 0; i < n_prefix; ++i) key[i] = seed_prefix[i];
        MT19937 g;
        for (int m = lo; m < hi; ++m) {
            key[n_prefix] = (uint32_t)(member0 + m);
            g.init_by_array(key, n_prefix + 1);
            bool has = false;
            double cached = 0.0;
            for (int i = 0; i < n_gauss; ++i) {
                double val;
                if (has) {
                    val = cached;
                    has = false;
                } else {
                    double x1, x2, r2;
                    do {
                        x1 = 2.0 * g.next_double() - 1.0;
                        x2 = 2.0 * g.next_double() - 1.0;
                        r2 = x1 * x1 + x2 * x2;
                    } while (r2 >= 1.0 || r2 == 0.0);
                    double f = sqrt(-2.0 * log(r2) / r2);
                    cached = f * x1;
                    has = true;
                    val = f * x2;
                }
                gauss_out[(size_t)m * n_gauss + i] = val;
            }
            for (int i = 0; i < n_uniform; ++i) uniform_out[(size_t)m * n_uniform + i] = g.next_double();
        }
    };
    if (n_threads <= 1) {
        work(0, n_members);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) {
            int lo = (int)((int64_t)n_members * t / n_threads), hi = (int)((int64_t)n_members * (t + 1) / n_threads);
            pool.emplace_back(work, lo, hi);
        }
        for (auto &th : pool) th.join();
    }
    return QSX_OK;
}

// ------------------------------------------- seeded disorder streams (device)
// The same streams generated on the GPU, one thread per ensemble member (the
// Mersenne-Twister state lives in the thread's local memory): integer arithmetic and
// the uniform doubles are bit-identical with the host replay; the polar Box-Muller
// transform uses the device log/sqrt (<= 1 ulp from the host libm).  Writes
// out[m][i] = scale * randn_i of member member0 + m.  Removes the host replay and the
// H2D copy from the end-to-end path of a disorder ensemble.
// The serial part of the seeding (init_by_array: 1247 dependent steps after the member-independent
// init_genrand(19650218), whose 624 words come from a table) carries the previous word in a register
// and loads the words it mixes in eight at a time ahead of the chain, so a step is a handful of
// integer operations instead of a store -> load round trip through local memory (0.33 -> see
// profiles/README.md ms per 1e4 members); the first outputs are generated straight from the seeded
// state (word k of the first twist needs words k, k + 1 and k + 397 of it), the full twist only
// runs if a member needs more than 227 words.
__device__ uint32_t qsx_mt_seed_table[624];

__global__ void __launch_bounds__(64) sample_streams_kernel(const uint32_t *prefix, int n_prefix, long long member0,
                                                            int n_members, int n_gauss, double scale, double *out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_members) return;
    uint32_t key[16];
    for (int i = 0; i < n_prefix; ++i) key[i] = prefix[i];
    key[n_prefix] = (uint32_t)(member0 + m);
    const int len = n_prefix + 1;
    uint32_t mt[624];
    // ---- init_by_array, first loop: i = 1 .. 623, then the wrapped step at i = 1
    uint32_t prev = qsx_mt_seed_table[0];
    int j = 0;
    for (int i = 1; i < 624; i += 8) {
        uint32_t old[8], add[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            old[u] = i + u < 624 ? qsx_mt_seed_table[i + u] : 0U;
            add[u] = key[j] + (uint32_t)j;
            if (++j >= len) j = 0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i + u < 624) {
                prev = (old[u] ^ ((prev ^ (prev >> 30)) * 1664525U)) + add[u];
                mt[i + u] = prev;
            }
        }
    }
    // 623 steps done, j advanced 624 times (one too many: the last group is padded): realign
    j = 623 % len;
    prev = (mt[1] ^ ((prev ^ (prev >> 30)) * 1664525U)) + key[j] + (uint32_t)j;      // mt[0] = mt[623]; i = 1
    mt[1] = prev;
    // ---- second loop: i = 2 .. 623, then the wrapped step at i = 1
    for (int i = 2; i < 624; i += 8) {
        uint32_t old[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) old[u] = i + u < 624 ? mt[i + u] : 0U;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i + u < 624) {
                prev = (old[u] ^ ((prev ^ (prev >> 30)) * 1566083941U)) - (uint32_t)(i + u);
                mt[i + u] = prev;
            }
        }
    }
    prev = (mt[1] ^ ((prev ^ (prev >> 30)) * 1566083941U)) - 1U;                        // mt[0] = mt[623]; i = 1
    mt[1] = prev;
    mt[0] = 0x80000000U;
    // ---- outputs
    int pos = 0;
    bool twisted = false;
    auto next32 = [&]() -> uint32_t {
        uint32_t y;
        if (!twisted && pos < 227) {
            const uint32_t w = (mt[pos] & 0x80000000U) | (mt[pos + 1] & 0x7fffffffU);
            y = mt[pos + 397] ^ (w >> 1) ^ ((w & 1U) ? 0x9908b0dfU : 0U);
            ++pos;
        } else {
            if (!twisted || pos >= 624) {
                // full twist of the current state (the direct outputs above did not modify it)
                for (int k = 0; k < 624; ++k) {
                    const uint32_t w = (mt[k] & 0x80000000U) | (mt[(k + 1) % 624] & 0x7fffffffU);
                    mt[k] = mt[(k + 397) % 624] ^ (w >> 1) ^ ((w & 1U) ? 0x9908b0dfU : 0U);
                }
                if (twisted) pos = 0;
                twisted = true;
            }
            y = mt[pos++];
        }
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    };
    auto next_double = [&]() -> double {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    };
    bool has = false;
    double cached = 0.0;
    for (int i = 0; i < n_gauss; ++i) {
        double val;
        if (has) {
            val = cached;
            has = false;
        } else {
            double x1, x2, r2;
            do {
                x1 = 2.0 * next_double() - 1.0;
                x2 = 2.0 * next_double() - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            double f = sqrt(-2.0 * log(r2) / r2);
            cached = f * x1;
            has = true;
            val = f * x2;
        }
        out[(size_t)m * n_gauss + i] = scale * val;
    }
}

extern "C" int qsx_sample_gauss_device(const uint32_t *seed_prefix, int32_t n_prefix, int64_t member0,
                                       int32_t n_members, int32_t n_gauss, double scale, void *out_dev,
                                       void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    QSX_REQUIRE(n_prefix >= 0 && n_prefix < 15 && n_members > 0 && n_gauss > 0 && out_dev,
                "qsx_sample_gauss_device: bad arguments");
    QSX_REQUIRE(member0 >= 0 && member0 + n_members <= (int64_t)0xffffffffLL, "member index out of range");
    {
        // the table is a per-device symbol: upload it once for every device this process uses
        static std::mutex seed_mu;
        static bool seeded[64] = {false};
        int dev = 0;
        QSX_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(seed_mu);
        if (dev < 0 || dev >= 64 || !seeded[dev]) {
            MT19937 g;
            g.init_genrand(19650218U);
            QSX_CUDA(cudaMemcpyToSymbolAsync(qsx_mt_seed_table, g.mt, sizeof(g.mt), 0, cudaMemcpyHostToDevice, stream));
            QSX_CUDA(cudaStreamSynchronize(stream));        // g.mt lives on this stack frame
            if (dev >= 0 && dev < 64) seeded[dev] = true;
        }
    }
    DevBuf<uint32_t> prefix;
    if (n_prefix > 0) QSX_CUDA(prefix.upload(seed_prefix, (size_t)n_prefix, stream));
    sample_streams_kernel<<<(n_members + 63) / 64, 64, 0, stream>>>(prefix.p, n_prefix, member0, n_members, n_gauss,
                                                                    scale, (double *)out_dev);
    qsx_launch_counter += 1;
    QSX_CUDA(cudaGetLastError());
    // no host synchronisation: the result stays on the device and `prefix` returns to the
    // stream-ordered scratch pool (all calls of a thread share one stream)
    return QSX_OK;
}
