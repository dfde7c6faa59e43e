// snapb200.cu -- C ABI of the B200 alignment core (include/snapb200.h): index residency, scratch tiers,
// launches, pinned double-buffered batch pipeline, statistics.  Host side of the library; all alignment
// arithmetic is in the kernels (kernels.cuh and the headers it includes).  There is no CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "index_build.cuh"
#include "characterize.cuh"  // after kernels.cuh: uses DevBatch, Counters, fetch_work
#include "iokernels.cuh"     // FASTQ / SAM batch kernels (row f2); uses lv_cigar_warp, stage_window
#include "bgzf_kernels.cuh"  // BGZF blocks (row f4b)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

thread_local char g_last_error[512] = "";

static std::mutex g_slot_mutex;
static bool g_slot_used[64][MAX_INDEX_SLOTS];

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *snapb200_last_error(void) { return g_last_error; }
extern "C" int snapb200_abi_version(void) { return SNAPB200_ABI_VERSION; }
extern "C" int snapb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---- small RAII-free helpers ------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, bool zero_new = false)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        if (zero_new) cudaMemset(p, 0, want);
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// pow(double,int) as the reference's gnu++98 build evaluates it (libstdc++ std::pow(double,int) ==
// __builtin_powi == libgcc __powidf2), used for pow(1 - SNP_PROB, seedLen) (BaseAligner.cpp:1227)
static double powi_ref(double x, int m)
{
    unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
    double y = (n % 2) ? x : 1;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? 1 / y : y;
}

// computeMAPQ with libm, for the device's rare "too close to an integer to call" requests (mapq.h:32-65)
static int compute_mapq_host(double p_all, double p_best, int score, int popular)
{
    if (!(p_all > p_best)) p_all = p_best;
    if (p_all == p_best && popular == 0 && score < 5) return 70;
    double correct = p_best / p_all;
    int base;
    if (correct >= 1) base = 69;
    else {
        int v = (int)(-10 * log10(1 - correct));
        base = v < 69 ? v : 69;
    }
    int pen = popular - 10;
    if (pen < 0) pen = 0;
    base -= pen / 2;
    return base > 0 ? base : 0;
}

#define FIX_CAP 65536

// Internal sessions of an index handle: how many callers of the synchronous batch entry points (and batches of the RNA pipeline) can
// have their device work in flight at once; a further caller waits for one.  Each owns its scratch tiers (a few hundred MB on C3).
#define MAX_BATCH_SESSIONS 4
static int batch_sessions()
{
    static const int n = [] {
        const char *e = getenv("SNAPB200_SESSIONS");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= MAX_BATCH_SESSIONS ? v : 2;
    }();
    return n;
}

static bool turns_enabled()
{
    static const bool on = getenv("SNAPB200_NO_TURNS") == nullptr;  // SNAPB200_NO_TURNS=1: measurement, the behaviour before
    return on;
}

struct snapb200_index {
    int device = 0;
    int slot = -1;  // position of `dev` in this device's c_index[]
    int sm_count = 0;
    DevIndex dev;
    snapb200_index_info info;
    std::vector<void *> allocs;
    std::vector<std::string> piece_names;
    const char *sam_names_blob = nullptr;     // piece names in HBM, uploaded by the first snapb200_sam_batch call
    const uint32_t *sam_names_off = nullptr;
    std::vector<uint64_t> table_sizes, table_used;
    cudaStream_t stream = nullptr;
    // scratch shared by the synchronous batch entry points (sessions own theirs)
    unsigned long long *stats = nullptr;  // SNAPB200_STATS_WORDS counters in HBM
    // The synchronous *_batch entry points run on one of two internal sessions (own stream + device buffers each), so two
    // host threads (the reference runs -t N of them) can have batches in flight at once: the tail of one batch's kernels
    // overlaps the head of the other's.  A third concurrent caller waits.
    struct snapb200_session *batch_session[MAX_BATCH_SESSIONS] = {nullptr, nullptr, nullptr, nullptr};
    std::mutex batch_mutex[MAX_BATCH_SESSIONS];
    std::mutex run_turn;  // see paired_chunks: the kernels of one chunk at a time, the copies of the other session overlap them
    std::atomic<unsigned> batch_rr{0};
};

static int upload(snapb200_index *x, const void *src, size_t bytes, void **dst, size_t pad_before = 0, size_t pad_after = 0, int pad_byte = 0)
{
    void *p = nullptr;
    size_t total = bytes + pad_before + pad_after;
    if (total == 0) total = 16;
    cudaError_t e = cudaMalloc(&p, total);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
    x->allocs.push_back(p);
    x->info.device_bytes += total;
    if (pad_before || pad_after) CUDA_TRY(cudaMemset(p, pad_byte, total));
    if (bytes) CUDA_TRY(cudaMemcpy((char *)p + pad_before, src, bytes, cudaMemcpyHostToDevice));
    *dst = (char *)p + pad_before;
    return 0;
}

static int finish_index(snapb200_index *x)
{
    // probability tables, computed on the host with libm exactly like LandauVishkin.cpp:601-653
    const double snp = 0.001, gap_open = 0.001, gap_extend = 0.5;
    std::vector<double> phred(256), indel(64), perfect(SNAPB200_MAX_READ_LENGTH + 1);
    indel[0] = 1.0;
    indel[1] = gap_open;
    for (int i = 2; i < 64; i++) indel[i] = indel[i - 1] * gap_extend;
    for (int i = 0; i < 256; i++) phred[i] = snp;
    for (int i = 33; i <= 93 + 33; i++) phred[i] = 1.0 - (1.0 - pow(10.0, -1.0 * (i - 33.0) / 10.0)) * (1.0 - snp);
    perfect[0] = 1.0;
    for (int i = 1; i <= SNAPB200_MAX_READ_LENGTH; i++) perfect[i] = perfect[i - 1] * (1 - snp);
    void *p;
    int rc;
    if ((rc = upload(x, phred.data(), phred.size() * 8, &p))) return rc;
    x->dev.phred = (const double *)p;
    if ((rc = upload(x, indel.data(), indel.size() * 8, &p))) return rc;
    x->dev.indel = (const double *)p;
    if ((rc = upload(x, perfect.data(), perfect.size() * 8, &p))) return rc;
    x->dev.perfect = (const double *)p;
    x->dev.seed_prob = powi_ref(1 - snp, (int)x->dev.seed_len);
    CUDA_TRY(cudaMalloc((void **)&x->stats, SNAPB200_STATS_WORDS * 8));
    CUDA_TRY(cudaMemset(x->stats, 0, SNAPB200_STATS_WORDS * 8));
    CUDA_TRY(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, x->device));
    x->sm_count = prop.multiProcessorCount;
    x->info.device = x->device;
    // publish the descriptor in constant memory
    {
        std::lock_guard<std::mutex> g(g_slot_mutex);
        if (x->device < 0 || x->device >= 64) return set_error(SNAPB200_ERR_ARG, "device %d out of range", x->device);
        for (int q = 0; q < MAX_INDEX_SLOTS && x->slot < 0; q++) if (!g_slot_used[x->device][q]) { g_slot_used[x->device][q] = true; x->slot = q; }
        if (x->slot < 0) return set_error(SNAPB200_ERR_ARG, "more than %d indices open on device %d", MAX_INDEX_SLOTS, x->device);
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_index, &x->dev, sizeof(DevIndex), (size_t)x->slot * sizeof(DevIndex)));
    return 0;
}

// Host threads that wait for the device sleep instead of spinning (cudaDeviceScheduleBlockingSync).  The reference runs as many worker
// threads as the box has cores; with the default policy every thread inside a cudaStreamSynchronize burns a core that a thread
// replaying GTF counters or formatting SAM needs, and is itself descheduled for whole time slices between the ~40 host round trips
// of a batch (measured through the command line: 145 ms per 32 k-pair batch against 28 ms when the cores are free).
// SNAPB200_SPIN=1 keeps the driver's default.  Harmless if the context already exists with other flags (e.g. created by PyTorch).
static void prefer_blocking_sync(int device)
{
    static std::mutex m;
    static bool done[64];
    std::lock_guard<std::mutex> g(m);
    if (device < 0 || device >= 64 || done[device]) return;
    done[device] = true;
    if (getenv("SNAPB200_SPIN")) return;
    if (cudaSetDevice(device) == cudaSuccess && cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) cudaGetLastError();
}

static int make_index(int device, uint32_t seed_len, uint32_t padding, uint32_t n_tables, const uint64_t *table_sizes,
                      const void *tables, const uint32_t *overflow, uint32_t overflow_words, const uint8_t *bases,
                      uint32_t n_bases, const uint32_t *piece_offsets, uint32_t n_pieces, snapb200_index **out,
                      void *adopt_tables = nullptr, void *adopt_overflow = nullptr)
{
    if (n_bases > 0xffffff00u) return set_error(SNAPB200_ERR_ARG, "genome of %u bases: locations must stay 256 below 2^32 (the reference stops at 0xfffffff0, GenomeIndex.cpp:372)", n_bases);
    // adopt_*: device allocations (from the device-side builder) to take over instead of uploading host copies
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_error(SNAPB200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_error(SNAPB200_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    if (seed_len < 16 || seed_len > 25) return set_error(SNAPB200_ERR_ARG, "seed length %u unsupported (16..25, SeedSequencer.h)", seed_len);
    uint32_t expect_tables = 1;
    for (uint32_t i = 16; i < seed_len; i++) expect_tables *= 4;
    if (n_tables != expect_tables) return set_error(SNAPB200_ERR_IO, "index has %u hash tables, seed length %u needs %u", n_tables, seed_len, expect_tables);
    prefer_blocking_sync(device);
    CUDA_TRY(cudaSetDevice(device));
    snapb200_index *x = new snapb200_index();
    x->device = device;
    memset(&x->dev, 0, sizeof(x->dev));
    memset(&x->info, 0, sizeof(x->info));
    std::vector<uint64_t> start(n_tables);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_tables; i++) { start[i] = total; total += table_sizes[i]; }
    void *p;
    int rc = 0;
    do {
        if (adopt_tables) {
            x->allocs.push_back(adopt_tables);
            x->info.device_bytes += total * sizeof(HtEntry);
            p = adopt_tables;
        } else if ((rc = upload(x, tables, total * sizeof(HtEntry), &p))) break;
        x->dev.tables = (const HtEntry *)p;
        if ((rc = upload(x, start.data(), n_tables * 8, &p))) break;
        x->dev.table_start = (const uint64_t *)p;
        if ((rc = upload(x, table_sizes, n_tables * 8, &p))) break;
        x->dev.table_size = (const uint64_t *)p;
        if (adopt_overflow) {
            x->allocs.push_back(adopt_overflow);
            x->info.device_bytes += (size_t)overflow_words * 4;
            p = adopt_overflow;
        } else if ((rc = upload(x, overflow, (size_t)overflow_words * 4, &p, 0, 16))) break;
        x->dev.overflow = (const uint32_t *)p;
        if ((rc = upload(x, bases, n_bases, &p, GENOME_PAD, GENOME_PAD, 'n'))) break;
        x->dev.genome = (const uint8_t *)p;
        if ((rc = upload(x, piece_offsets, (size_t)n_pieces * 4, &p))) break;
        x->dev.piece_begin = (const uint32_t *)p;
        x->dev.n_bases = n_bases; x->dev.n_pieces = n_pieces; x->dev.seed_len = seed_len; x->dev.n_tables = n_tables;
        x->dev.padding = padding;
        x->info.n_bases = n_bases; x->info.n_pieces = n_pieces; x->info.seed_len = seed_len; x->info.n_hash_tables = n_tables;
        x->info.overflow_table_size = overflow_words; x->info.chromosome_padding = padding; x->info.hash_table_entries = total;
        x->table_sizes.assign(table_sizes, table_sizes + n_tables);
        rc = finish_index(x);
    } while (0);
    if (rc) { snapb200_index_close(x); return rc; }
    *out = x;
    return 0;
}

extern "C" int snapb200_index_from_memory(int device, uint32_t seed_len, uint32_t chromosome_padding, uint32_t n_hash_tables,
                                          const uint64_t *table_sizes, const void *tables, const uint32_t *overflow,
                                          uint32_t overflow_words, const uint8_t *bases, uint32_t n_bases,
                                          const uint32_t *piece_offsets, uint32_t n_pieces, snapb200_index **out)
{
    if (!out || !table_sizes || !tables || !bases) return set_error(SNAPB200_ERR_ARG, "null argument");
    return make_index(device, seed_len, chromosome_padding, n_hash_tables, table_sizes, tables, overflow, overflow_words, bases,
                      n_bases, piece_offsets, n_pieces, out);
}

static bool read_file(const std::string &path, std::vector<char> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)sz);
    size_t got = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz;
}

// A read-only mapping of a whole file: the index files go from the page cache to HBM without a copy in host memory (at 3.1 Gbp the
// hash tables are 48 GB).
struct MappedFile {
    const char *p = nullptr;
    size_t size = 0;
    bool open(const std::string &path)
    {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        size = (size_t)st.st_size;
        if (size) {
            void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); return false; }
            madvise(m, size, MADV_SEQUENTIAL);
            p = (const char *)m;
        }
        ::close(fd);
        return true;
    }
    ~MappedFile() { if (p) munmap((void *)p, size); }
};

// File formats: GenomeIndex.cpp:646-710 (GenomeIndex, OverflowTable, GenomeIndexHash), HashTable.cpp:181-215
// (per table: u32 magic 0xb111b010, size_t tableSize, size_t usedElementCount, entries), Genome.cpp:126-158.
extern "C" int snapb200_index_open(const char *dir, int device, snapb200_index **out)
{
    if (!dir || !out) return set_error(SNAPB200_ERR_ARG, "null argument");
    std::string d(dir);
    std::vector<char> meta;
    if (!read_file(d + "/GenomeIndex", meta)) return set_error(SNAPB200_ERR_IO, "cannot read %s/GenomeIndex", dir);
    meta.push_back(0);
    unsigned major, minor, n_tables, overflow_words, seed_len, padding;
    if (sscanf(meta.data(), "%u %u %u %u %u %u", &major, &minor, &n_tables, &overflow_words, &seed_len, &padding) != 6)
        return set_error(SNAPB200_ERR_IO, "%s/GenomeIndex: expected six integers", dir);
    MappedFile ovf, hash, genome;
    if (!ovf.open(d + "/OverflowTable") || ovf.size < (size_t)overflow_words * 4)
        return set_error(SNAPB200_ERR_IO, "cannot read %s/OverflowTable (%u words expected)", dir, overflow_words);
    if (!hash.open(d + "/GenomeIndexHash")) return set_error(SNAPB200_ERR_IO, "cannot read %s/GenomeIndexHash", dir);
    std::vector<uint64_t> sizes(n_tables), file_pos(n_tables);
    size_t pos = 0;
    uint64_t total = 0;
    for (unsigned i = 0; i < n_tables; i++) {
        if (pos + 20 > hash.size) return set_error(SNAPB200_ERR_IO, "GenomeIndexHash truncated at table %u", i);
        uint32_t magic;
        uint64_t size, used;
        memcpy(&magic, hash.p + pos, 4); memcpy(&size, hash.p + pos + 4, 8); memcpy(&used, hash.p + pos + 12, 8);
        pos += 20;
        if (magic != 0xb111b010u) return set_error(SNAPB200_ERR_IO, "GenomeIndexHash: bad magic at table %u", i);
        if (size == 0 || pos + size * 12 > hash.size) return set_error(SNAPB200_ERR_IO, "GenomeIndexHash: bad size at table %u", i);
        sizes[i] = size;
        file_pos[i] = pos;
        total += size;
        pos += size * 12;
    }
    if (!genome.open(d + "/Genome") || !genome.size) return set_error(SNAPB200_ERR_IO, "cannot read %s/Genome", dir);
    unsigned n_bases = 0, n_pieces = 0;
    size_t gp = 0;
    const char *gbeg = genome.p, *gend = genome.p + genome.size;
    {
        const char *eol = (const char *)memchr(gbeg, '\n', genome.size);
        if (!eol) return set_error(SNAPB200_ERR_IO, "Genome: no header line");
        std::string line(gbeg, eol);
        if (sscanf(line.c_str(), "%u %u", &n_bases, &n_pieces) != 2) return set_error(SNAPB200_ERR_IO, "Genome: bad header");
        gp = (size_t)(eol - gbeg) + 1;
    }
    std::vector<uint32_t> pieces(n_pieces);
    std::vector<std::string> names;
    for (unsigned i = 0; i < n_pieces; i++) {
        const char *eol = gp < genome.size ? (const char *)memchr(gbeg + gp, '\n', genome.size - gp) : nullptr;
        if (!eol) return set_error(SNAPB200_ERR_IO, "Genome: truncated piece table");
        std::string line(gbeg + gp, eol);
        pieces[i] = (uint32_t)atoi(line.c_str());
        size_t sp = line.find(' ');
        names.push_back(sp == std::string::npos ? std::string("piece") + std::to_string(i) : line.substr(sp + 1));
        gp = (size_t)(eol - gbeg) + 1;
    }
    if (gp + n_bases > genome.size) return set_error(SNAPB200_ERR_IO, "Genome: %u bases expected", n_bases);
    (void)gend;
    // the tables go table by table from the mapping to their place in one device allocation, which make_index adopts (and frees
    // if it fails after adopting; what it rejects before that is checked here first)
    if (seed_len < 16 || seed_len > 25) return set_error(SNAPB200_ERR_ARG, "seed length %u unsupported (16..25, SeedSequencer.h)", seed_len);
    {
        uint32_t expect_tables = 1;
        for (uint32_t i = 16; i < seed_len; i++) expect_tables *= 4;
        if (n_tables != expect_tables) return set_error(SNAPB200_ERR_IO, "index has %u hash tables, seed length %u needs %u", n_tables, seed_len, expect_tables);
    }
    if (n_bases > 0xffffff00u) return set_error(SNAPB200_ERR_ARG, "genome of %u bases: locations must stay 256 below 2^32 (the reference stops at 0xfffffff0, GenomeIndex.cpp:372)", n_bases);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(SNAPB200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_error(SNAPB200_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    prefer_blocking_sync(device);
    CUDA_TRY(cudaSetDevice(device));
    HtEntry *d_tables = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_tables, std::max<uint64_t>(total, 1) * sizeof(HtEntry));
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%llu) for the hash tables failed: %s", (unsigned long long)(total * sizeof(HtEntry)), cudaGetErrorString(e));
    uint64_t start = 0;
    for (unsigned i = 0; i < n_tables; i++) {
        e = cudaMemcpy(d_tables + start, hash.p + file_pos[i], sizes[i] * sizeof(HtEntry), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d_tables); return set_error(SNAPB200_ERR_CUDA, "uploading hash table %u: %s", i, cudaGetErrorString(e)); }
        start += sizes[i];
    }
    int rc = make_index(device, seed_len, padding, n_tables, sizes.data(), nullptr, (const uint32_t *)ovf.p, overflow_words,
                        (const uint8_t *)genome.p + gp, n_bases, pieces.data(), n_pieces, out, d_tables, nullptr);
    if (!rc) (*out)->piece_names = names;
    return rc;
}

// ---- device-side index construction (index_build.cuh) -----------------------------------------------------------
template <class T>
static int dev_alloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return 0;
}

extern "C" int snapb200_index_build(int device, const uint8_t *bases, uint32_t n_bases, const uint32_t *piece_offsets,
                                    const char *const *piece_names, uint32_t n_pieces, uint32_t seed_len, uint32_t chromosome_padding,
                                    double slack, snapb200_index **out)
{
    if (!bases || !out || (n_pieces && !piece_offsets)) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (seed_len < 16 || seed_len > 25) return set_error(SNAPB200_ERR_ARG, "seed length %u unsupported (16..25)", seed_len);
    if (n_bases > 0xfffffff0u || n_bases <= seed_len + 1) return set_error(SNAPB200_ERR_ARG, "genome size %u out of range", n_bases);
    if (!(slack >= 0.05)) slack = 0.3;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(SNAPB200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_error(SNAPB200_ERR_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    // the reference indexes locations [0, nBases - seedLen - 1) (GenomeIndex.cpp:455-470)
    const uint32_t n_pos = n_bases - seed_len - 1;
    uint32_t n_tables = 1;
    for (uint32_t i = 16; i < seed_len; i++) n_tables *= 4;
    uint8_t *d_genome = nullptr;
    unsigned long long *k0 = nullptr, *k1 = nullptr, *d_nvalid = nullptr, *d_tcount = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr, *head = nullptr, *rid = nullptr, *run_start = nullptr, *need = nullptr, *ovf_off = nullptr, *d_overflow = nullptr;
    void *tmp = nullptr;
    HtEntry *d_tables = nullptr;
    uint64_t *d_tstart = nullptr, *d_tsize = nullptr;
    int rc = 0;
    std::vector<uint64_t> sizes(n_tables), starts(n_tables), counts(n_tables);
    uint32_t overflow_words = 0;
    // every failure inside the block leaves through the cleanup below (a CUDA_TRY would return and leak up to ~120 GB of build buffers)
#define IB_TRY(expr)                                                                                                          \
    {                                                                                                                         \
        cudaError_t _e = (expr);                                                                                              \
        if (_e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); break; } \
    }
    do {
        if ((rc = dev_alloc(&d_genome, (size_t)n_bases + 64))) break;
        IB_TRY(cudaMemset(d_genome, 'n', (size_t)n_bases + 64));
        IB_TRY(cudaMemcpy(d_genome, bases, n_bases, cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(&k0, n_pos)) || (rc = dev_alloc(&k1, n_pos)) || (rc = dev_alloc(&v0, n_pos)) || (rc = dev_alloc(&v1, n_pos)) ||
            (rc = dev_alloc(&d_nvalid, 1)) || (rc = dev_alloc(&d_tcount, n_tables))) break;
        IB_TRY(cudaMemset(d_nvalid, 0, 8));
        IB_TRY(cudaMemset(d_tcount, 0, (size_t)n_tables * 8));
        const int T = 256;
        ib_emit_kernel<<<(n_pos + T - 1) / T, T>>>(d_genome, n_pos, seed_len, k0, v0, d_nvalid);
        IB_TRY(cudaGetLastError());
        size_t tmp_bytes = 0;
        const int end_bit = 2 * (int)seed_len + 1 > 63 ? 64 : 64;  // invalid keys (all ones) must sort last: use all bits
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, (unsigned long long)n_pos, 0, end_bit);
        IB_TRY(cudaMalloc(&tmp, tmp_bytes));
        IB_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (unsigned long long)n_pos, 0, end_bit));
        unsigned long long n_valid64 = 0;
        IB_TRY(cudaMemcpy(&n_valid64, d_nvalid, 8, cudaMemcpyDeviceToHost));
        const uint32_t n_valid = (uint32_t)n_valid64;
        cudaFree(tmp); tmp = nullptr;
        cudaFree(k0); k0 = nullptr;
        cudaFree(v0); v0 = nullptr;
        uint32_t n_runs = 0;
        if (n_valid) {
            if ((rc = dev_alloc(&head, n_valid)) || (rc = dev_alloc(&rid, n_valid))) break;
            ib_heads_kernel<<<(n_valid + T - 1) / T, T>>>(k1, n_valid, head);
            tmp_bytes = 0;
            cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, head, rid, (unsigned long long)n_valid);
            IB_TRY(cudaMalloc(&tmp, tmp_bytes));
            IB_TRY(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, rid, (unsigned long long)n_valid));
            cudaFree(tmp); tmp = nullptr;
            IB_TRY(cudaMemcpy(&n_runs, rid + (n_valid - 1), 4, cudaMemcpyDeviceToHost));
            // buffers are released as soon as they are dead: at 3.1 Gbp the peak stays near 120 GB of the 180 GB
            if ((rc = dev_alloc(&run_start, (size_t)n_runs + 1))) break;
            ib_run_start_kernel<<<(n_valid + T - 1) / T, T>>>(head, rid, n_valid, run_start);
            IB_TRY(cudaMemcpy(run_start + n_runs, &n_valid, 4, cudaMemcpyHostToDevice));
            cudaFree(head); head = nullptr;
            if ((rc = dev_alloc(&need, (size_t)n_runs + 1)) || (rc = dev_alloc(&ovf_off, (size_t)n_runs + 1))) break;
            IB_TRY(cudaMemset(d_nvalid, 0, 8));
            ib_run_need_kernel<<<(n_runs + T - 1) / T, T>>>(run_start, n_runs, need, d_nvalid);
            IB_TRY(cudaMemset(need + n_runs, 0, 4));
            unsigned long long need_total = 0;  // 64-bit total first: the 32-bit prefix sum below must not wrap
            IB_TRY(cudaMemcpy(&need_total, d_nvalid, 8, cudaMemcpyDeviceToHost));
            if ((uint64_t)n_bases + need_total > 0xfffffff0ull) { rc = set_error(SNAPB200_ERR_LIMIT, "too many overflow entries for this seed length (GenomeIndex.cpp:492-495)"); break; }
            tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, need, ovf_off, (unsigned long long)n_runs + 1);
            IB_TRY(cudaMalloc(&tmp, tmp_bytes));
            IB_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, need, ovf_off, (unsigned long long)n_runs + 1));
            cudaFree(tmp); tmp = nullptr;
            cudaFree(need); need = nullptr;
            IB_TRY(cudaMemcpy(&overflow_words, ovf_off + n_runs, 4, cudaMemcpyDeviceToHost));
            if ((rc = dev_alloc(&d_overflow, (size_t)overflow_words + 4))) break;
            ib_fill_overflow_kernel<<<(n_valid + T - 1) / T, T>>>(rid, run_start, ovf_off, v1, n_valid, d_overflow);
            IB_TRY(cudaGetLastError());
            IB_TRY(cudaDeviceSynchronize());
            cudaFree(rid); rid = nullptr;
            ib_count_tables_kernel<<<(n_runs + T - 1) / T, T>>>(k1, run_start, n_runs, d_tcount);
            IB_TRY(cudaGetLastError());
        }
        IB_TRY(cudaMemcpy(counts.data(), d_tcount, (size_t)n_tables * 8, cudaMemcpyDeviceToHost));
        uint64_t total = 0;
        for (uint32_t i = 0; i < n_tables; i++) {
            sizes[i] = std::max<uint64_t>(64, (uint64_t)((double)counts[i] * (1.0 + slack) * 1.1) + 16);
            starts[i] = total;
            total += sizes[i];
        }
        if ((rc = dev_alloc(&d_tables, total)) || (rc = dev_alloc(&d_tstart, n_tables)) || (rc = dev_alloc(&d_tsize, n_tables))) break;
        IB_TRY(cudaMemset(d_tables, 0xff, total * sizeof(HtEntry)));  // free entries: value1 == InvalidGenomeLocation
        IB_TRY(cudaMemcpy(d_tstart, starts.data(), (size_t)n_tables * 8, cudaMemcpyHostToDevice));
        IB_TRY(cudaMemcpy(d_tsize, sizes.data(), (size_t)n_tables * 8, cudaMemcpyHostToDevice));
        if (n_runs) {
            ib_insert_kernel<<<(n_runs + T - 1) / T, T>>>(k1, run_start, ovf_off, v1, n_runs, n_bases, d_tables, d_tstart, d_tsize);
            IB_TRY(cudaGetLastError());
        }
        IB_TRY(cudaDeviceSynchronize());
    } while (0);
#undef IB_TRY
    if (!d_overflow && !rc) rc = dev_alloc(&d_overflow, 4);  // a genome without repeated seeds still gets a table to point at
    void *frees[] = {d_genome, k0, k1, d_nvalid, d_tcount, v0, v1, head, rid, run_start, need, ovf_off, tmp, d_tstart, d_tsize};
    for (void *p : frees) if (p) cudaFree(p);
    if (rc) {
        if (d_tables) cudaFree(d_tables);
        if (d_overflow) cudaFree(d_overflow);
        return rc;
    }
    // the tables and the overflow table stay where they were built; the common constructor adopts them
    rc = make_index(device, seed_len, chromosome_padding, n_tables, sizes.data(), nullptr, nullptr, overflow_words, bases, n_bases, piece_offsets,
                    n_pieces, out, d_tables, d_overflow);
    if (!rc) {
        (*out)->table_used = counts;
        for (uint32_t i = 0; i < n_pieces; i++) (*out)->piece_names.push_back(piece_names && piece_names[i] ? piece_names[i] : ("piece" + std::to_string(i)));
    }
    return rc;
}

extern "C" int snapb200_index_save(snapb200_index *x, const char *dir)
{
    if (!x || !dir) return set_error(SNAPB200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(x->device));
    mkdir(dir, 0777);
    std::string d(dir);
    FILE *f = fopen((d + "/GenomeIndex").c_str(), "w");
    if (!f) return set_error(SNAPB200_ERR_IO, "cannot write %s/GenomeIndex", dir);
    fprintf(f, "%d %d %d %d %d %d", 1, 0, (int)x->dev.n_tables, (int)x->info.overflow_table_size, (int)x->dev.seed_len, (int)x->dev.padding);
    fclose(f);
    std::vector<char> buf;
    buf.resize((size_t)x->info.overflow_table_size * 4);
    if (!buf.empty()) CUDA_TRY(cudaMemcpy(buf.data(), x->dev.overflow, buf.size(), cudaMemcpyDeviceToHost));
    f = fopen((d + "/OverflowTable").c_str(), "wb");
    if (!f) return set_error(SNAPB200_ERR_IO, "cannot write %s/OverflowTable", dir);
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
    f = fopen((d + "/GenomeIndexHash").c_str(), "wb");
    if (!f) return set_error(SNAPB200_ERR_IO, "cannot write %s/GenomeIndexHash", dir);
    uint64_t start = 0;
    for (uint32_t i = 0; i < x->dev.n_tables; i++) {
        const uint32_t magic = 0xb111b010u;
        uint64_t size = x->table_sizes[i], used = i < x->table_used.size() ? x->table_used[i] : 0;
        buf.resize(size * sizeof(HtEntry));
        CUDA_TRY(cudaMemcpy(buf.data(), x->dev.tables + start, buf.size(), cudaMemcpyDeviceToHost));
        fwrite(&magic, 4, 1, f); fwrite(&size, 8, 1, f); fwrite(&used, 8, 1, f);
        fwrite(buf.data(), 1, buf.size(), f);
        start += size;
    }
    fclose(f);
    f = fopen((d + "/Genome").c_str(), "wb");
    if (!f) return set_error(SNAPB200_ERR_IO, "cannot write %s/Genome", dir);
    fprintf(f, "%d %d\n", (int)x->dev.n_bases, (int)x->dev.n_pieces);
    std::vector<uint32_t> pieces(x->dev.n_pieces);
    if (x->dev.n_pieces) CUDA_TRY(cudaMemcpy(pieces.data(), x->dev.piece_begin, (size_t)x->dev.n_pieces * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < x->dev.n_pieces; i++) fprintf(f, "%d %s\n", (int)pieces[i], i < x->piece_names.size() ? x->piece_names[i].c_str() : "piece");
    buf.resize(x->dev.n_bases);
    CUDA_TRY(cudaMemcpy(buf.data(), x->dev.genome, buf.size(), cudaMemcpyDeviceToHost));
    fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
    return 0;
}

extern "C" int snapb200_index_info_get(const snapb200_index *idx, snapb200_index_info *info)
{
    if (!idx || !info) return set_error(SNAPB200_ERR_ARG, "null argument");
    *info = idx->info;
    return 0;
}

// ---- sessions ---------------------------------------------------------------------------------------------------
struct snapb200_session {
    snapb200_index *idx = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr;
    float main_ms = 0;
    bool main_pending = false;
    uint32_t max_items = 0, max_read_len = 0;
    // resident batch
    DevBuf offsets[2], bases[2], quals[2];
    uint32_t n[2] = {0, 0};
    uint32_t max_len_seen = 0;
    // results + lists
    DevBuf single_res, paired_res, fb_single_res, retry_list, fallback_list, fb_positions, fix, counters;
    DevBuf mh_counts, mh_locs, mh_rcs, mh_scores;
    // scratch tiers
    DevBuf s_pool, s_anchors, s_lists, s_epochs, s_hitc, s_hitl, s_hitr;
    DevBuf p_cands, p_mates, p_anchors, p_lane_tables, p_order;
    DevBuf w_keys[2], w_vals[2], w_tmp;  // work ordering of the paired path (weigh_kernel + radix sort)
    DevBuf retry_tmp;                    // sorted retry list of the next scratch tier (kept: cudaMalloc / cudaFree synchronise the device)
    DevBuf f_scratch, f_work;            // per-warp scratch of filter_warp_kernel when it runs on this session's stream (rna batches)
    uint32_t anchors_tsize = 0;   // table size the anchor buffer was zeroed for
    uint32_t anchors_warps = 0;
    // last run
    float last_ms = 0;
    uint32_t last_launches = 0;
    uint64_t total_launches = 0;
    uint32_t last_n = 0;
    int last_kind = 0;  // 1 single, 2 paired
    std::vector<MapqFix> host_fix;
    int limit_hit = 0;
};

static uint32_t next_pow2(uint32_t v)
{
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

extern "C" int snapb200_session_create(snapb200_index *idx, uint32_t max_items, uint32_t max_read_len, snapb200_session **out)
{
    if (!idx || !out) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (max_read_len > SNAPB200_MAX_READ_LENGTH) return set_error(SNAPB200_ERR_ARG, "max_read_len %u > MAX_READ_LENGTH", max_read_len);
    CUDA_TRY(cudaSetDevice(idx->device));
    snapb200_session *s = new snapb200_session();
    s->idx = idx;
    s->max_items = max_items;
    s->max_read_len = max_read_len;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreate(&s->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&s->ev1)) != cudaSuccess || (e = cudaEventCreate(&s->evm0)) != cudaSuccess || (e = cudaEventCreate(&s->evm1)) != cudaSuccess) {
        snapb200_session_destroy(s);
        return set_error(SNAPB200_ERR_CUDA, "session_create: %s", cudaGetErrorString(e));
    }
    int rc;
    if ((rc = s->counters.ensure(sizeof(Counters))) || (rc = s->fix.ensure(sizeof(MapqFix) * FIX_CAP))) { snapb200_session_destroy(s); return rc; }
    *out = s;
    return 0;
}

extern "C" void snapb200_session_destroy(snapb200_session *s)
{
    if (!s) return;
    cudaSetDevice(s->idx->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    DevBuf *all[] = {&s->offsets[0], &s->offsets[1], &s->bases[0], &s->bases[1], &s->quals[0], &s->quals[1], &s->single_res,
                     &s->paired_res, &s->fb_single_res, &s->retry_list, &s->fallback_list, &s->fb_positions, &s->fix, &s->counters,
                     &s->mh_counts, &s->mh_locs, &s->mh_rcs, &s->mh_scores, &s->s_pool, &s->s_anchors, &s->s_lists, &s->s_epochs,
                     &s->s_hitc, &s->s_hitl, &s->s_hitr, &s->p_cands, &s->p_mates, &s->p_anchors, &s->p_lane_tables, &s->p_order,
                     &s->w_keys[0], &s->w_keys[1], &s->w_vals[0], &s->w_vals[1], &s->w_tmp, &s->retry_tmp, &s->f_scratch, &s->f_work};
    for (DevBuf *b : all) b->release();
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evm0) cudaEventDestroy(s->evm0);
    if (s->evm1) cudaEventDestroy(s->evm1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" void snapb200_index_close(snapb200_index *x)
{
    if (!x) return;
    cudaSetDevice(x->device);
    for (int i = 0; i < MAX_BATCH_SESSIONS; i++) if (x->batch_session[i]) snapb200_session_destroy(x->batch_session[i]);
    for (void *p : x->allocs) cudaFree(p);
    if (x->slot >= 0) { std::lock_guard<std::mutex> g(g_slot_mutex); g_slot_used[x->device][x->slot] = false; }
    if (x->stats) cudaFree(x->stats);
    if (x->stream) cudaStreamDestroy(x->stream);
    delete x;
}

static int validate_batch(const snapb200_read_batch *b, uint32_t *max_len)
{
    if (!b || (b->n && (!b->offsets || !b->bases || !b->quals))) return set_error(SNAPB200_ERR_ARG, "null read batch");
    uint32_t m = 0;
    for (uint32_t i = 0; i < b->n; i++) {
        if (b->offsets[i + 1] < b->offsets[i]) return set_error(SNAPB200_ERR_ARG, "read offsets not monotonic at %u", i);
        m = std::max(m, b->offsets[i + 1] - b->offsets[i]);
    }
    // the reference aborts on reads longer than maxReadSize (BaseAligner.cpp:609-613)
    if (m > SNAPB200_MAX_READ_LENGTH) return set_error(SNAPB200_ERR_ARG, "read of %u bases exceeds MAX_READ_LENGTH %d", m, SNAPB200_MAX_READ_LENGTH);
    *max_len = m;
    return 0;
}

extern "C" int snapb200_session_upload(snapb200_session *s, int slot, const snapb200_read_batch *reads)
{
    if (!s || slot < 0 || slot > 1) return set_error(SNAPB200_ERR_ARG, "bad session/slot");
    uint32_t m = 0;
    int rc = validate_batch(reads, &m);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(s->idx->device));
    const uint32_t n = reads->n;
    const size_t nb = n ? reads->offsets[n] : 0;
    if ((rc = s->offsets[slot].ensure((size_t)(n + 1) * 4))) return rc;
    if ((rc = s->bases[slot].ensure(nb + 16))) return rc;
    if ((rc = s->quals[slot].ensure(nb + 16))) return rc;
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(s->offsets[slot].p, reads->offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s->stream));
        if (nb) {
            CUDA_TRY(cudaMemcpyAsync(s->bases[slot].p, reads->bases, nb, cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaMemcpyAsync(s->quals[slot].p, reads->quals, nb, cudaMemcpyHostToDevice, s->stream));
        }
    }
    s->n[slot] = n;
    if (slot == 0) s->max_len_seen = m; else s->max_len_seen = std::max(s->max_len_seen, m);
    return 0;
}

static DevBatch dev_batch(const snapb200_session *s, int slot)
{
    DevBatch b;
    b.offsets = s->offsets[slot].as<uint32_t>();
    b.bases = s->bases[slot].as<uint8_t>();
    b.quals = s->quals[slot].as<uint8_t>();
    b.n = s->n[slot];
    return b;
}

static int read_counters(snapb200_session *s, Counters *c)
{
    CUDA_TRY(cudaMemcpyAsync(c, s->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

static int reset_work(snapb200_session *s)
{
    CUDA_TRY(cudaMemsetAsync(s->counters.p, 0, sizeof(uint32_t), s->stream));  // Counters::work
    return 0;
}

// grid: as many CTAs as fit, a multiple of the SM count
template <class K>
static int grid_for(K kernel, size_t smem, int sm_count, int *ctas_per_sm)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CTA_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    *ctas_per_sm = per_sm;
    return per_sm * sm_count;
}

static const size_t SCRATCH_BUDGET = (size_t)24 << 30;  // HBM the scratch of one launch may take

// Heaviest items first: weigh_kernel + a stable 16-bit radix sort of the item indices.  *order = nullptr when the batch is too
// small to bother (or SNAPB200_NO_ORDER is set).  mates: 2 = pairs (both resident batches), 1 = single reads (batch 0).
static int work_order(snapb200_session *s, int mates, uint32_t n, uint32_t max_hits, const uint32_t **order)
{
    *order = nullptr;
    if (n < 4096 || getenv("SNAPB200_NO_ORDER")) return 0;
    snapb200_index *x = s->idx;
    int rc;
    for (int q = 0; q < 2; q++) if ((rc = s->w_keys[q].ensure((size_t)n * 4)) || (rc = s->w_vals[q].ensure((size_t)n * 4))) return rc;
    const unsigned blocks = (unsigned)(((size_t)WEIGH_SEEDS * mates * n + 255) / 256);
    if (mates == 2)
        weigh_kernel<2><<<blocks, 256, 0, s->stream>>>(x->dev, dev_batch(s, 0), dev_batch(s, 1), n, max_hits, s->w_keys[0].as<uint32_t>(), s->w_vals[0].as<uint32_t>());
    else
        weigh_kernel<1><<<blocks, 256, 0, s->stream>>>(x->dev, dev_batch(s, 0), dev_batch(s, 0), n, max_hits, s->w_keys[0].as<uint32_t>(), s->w_vals[0].as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    s->last_launches++;
    cub::DoubleBuffer<uint32_t> kb(s->w_keys[0].as<uint32_t>(), s->w_keys[1].as<uint32_t>());
    cub::DoubleBuffer<uint32_t> vb(s->w_vals[0].as<uint32_t>(), s->w_vals[1].as<uint32_t>());
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 16, s->stream);
    if ((rc = s->w_tmp.ensure(tmp_bytes))) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->w_tmp.p, tmp_bytes, kb, vb, (int)n, 0, 16, s->stream));
    s->last_launches += 2;  // histogram + onesweep
    *order = vb.Current();
    return 0;
}

struct SingleTier { uint32_t pool_cap, tsize; int grid; };

static int launch_single(snapb200_session *s, const SingleCfg &cfg_in, const SingleTier &tier, const DevBatch b[2], int two_batches,
                         const uint32_t *positions, uint32_t n_items, snapb200_single_result *results, int mapq_divisor)
{
    snapb200_index *x = s->idx;
    SingleArgs a;
    memset(&a, 0, sizeof(a));
    a.ix_slot = x->slot;
    a.cfg = cfg_in;
    a.cfg.pool_cap = tier.pool_cap;
    a.cfg.tmask = tier.tsize - 1;
    a.b[0] = b[0];
    a.b[1] = b[1];
    a.two_batches = two_batches;
    a.positions = positions;
    a.items = nullptr;
    a.n_items = n_items;
    a.results = results;
    a.mh_counts = s->mh_counts.as<int32_t>(); a.mh_locs = s->mh_locs.as<uint32_t>();
    a.mh_rcs = s->mh_rcs.as<uint8_t>(); a.mh_scores = s->mh_scores.as<int32_t>();
    const size_t warps = (size_t)tier.grid * WARPS_PER_CTA;
    int rc;
    if ((rc = s->s_pool.ensure(warps * tier.pool_cap * sizeof(Elem)))) return rc;
    // anchors carry an epoch tag and must start zeroed; re-zero when the geometry changes
    const size_t anchor_bytes = warps * 2 * tier.tsize * sizeof(int2);
    if (anchor_bytes > s->s_anchors.cap || s->anchors_tsize != tier.tsize || s->anchors_warps != warps) {
        if ((rc = s->s_anchors.ensure(anchor_bytes))) return rc;
        CUDA_TRY(cudaMemsetAsync(s->s_anchors.p, 0, anchor_bytes, s->stream));
        if ((rc = s->s_epochs.ensure(warps * 4))) return rc;
        CUDA_TRY(cudaMemsetAsync(s->s_epochs.p, 0, warps * 4, s->stream));
        s->anchors_tsize = tier.tsize;
        s->anchors_warps = (uint32_t)warps;
    }
    if ((rc = s->s_lists.ensure(warps * 2 * a.cfg.n_lists * sizeof(int)))) return rc;
    if (a.cfg.max_hits_to_get) {
        if ((rc = s->s_hitc.ensure(warps * MAXK * 4))) return rc;
        if ((rc = s->s_hitl.ensure(warps * MAXK * 512 * 4))) return rc;
        if ((rc = s->s_hitr.ensure(warps * MAXK * 512))) return rc;
        a.hit_count = s->s_hitc.as<uint32_t>(); a.hit_loc = s->s_hitl.as<uint32_t>(); a.hit_rc = s->s_hitr.as<uint8_t>();
    }
    a.pool = s->s_pool.as<Elem>(); a.anchors = s->s_anchors.as<int2>(); a.lists = s->s_lists.as<int>();
    a.epochs = s->s_epochs.as<uint32_t>();
    a.ctr = s->counters.as<Counters>(); a.retry_list = s->retry_list.as<uint32_t>();
    a.fix = s->fix.as<MapqFix>(); a.fix_cap = FIX_CAP;
    a.stats = x->stats;
    a.mapq_divisor = mapq_divisor;
    if ((rc = reset_work(s))) return rc;
    const size_t smem = single_warp_shared(a.cfg.rl) * WARPS_PER_CTA;
    const bool time_it = s->main_pending && mapq_divisor == 1;
    if (time_it) CUDA_TRY(cudaEventRecord(s->evm0, s->stream));
    const bool prof = getenv("SNAPB200_PROF") != nullptr;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (prof) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); cudaEventRecord(pe0, s->stream); }
    single_kernel<<<tier.grid, CTA_THREADS, smem, s->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    if (prof) {
        cudaEventRecord(pe1, s->stream);
        cudaEventSynchronize(pe1);
        float ms = 0;
        cudaEventElapsedTime(&ms, pe0, pe1);
        fprintf(stderr, "[snapb200 prof] single_kernel grid=%d items=%u pool_cap=%u divisor=%d: %.2f ms\n", tier.grid, n_items, tier.pool_cap, mapq_divisor, ms);
        cudaEventDestroy(pe0); cudaEventDestroy(pe1);
    }
    if (time_it) { CUDA_TRY(cudaEventRecord(s->evm1, s->stream)); s->main_pending = false; }
    s->last_launches++;
    return 0;
}

static int single_cfg_from(const snapb200_index *x, const snapb200_single_params *p, uint32_t max_len, SingleCfg *cfg)
{
    if (!p) return set_error(SNAPB200_ERR_ARG, "null params");
    if (p->max_k + p->extra_search_depth >= MAXK) return set_error(SNAPB200_ERR_ARG, "max_k + extra_search_depth must be < %d (SingleAligner.cpp:117-121)", MAXK);
    if (p->max_hits == 0) return set_error(SNAPB200_ERR_ARG, "max_hits must be > 0");
    memset(cfg, 0, sizeof(*cfg));
    cfg->max_hits = p->max_hits; cfg->max_k = p->max_k; cfg->num_seeds = p->num_seeds; cfg->extra = p->extra_search_depth;
    cfg->explore = p->explore_popular_seeds; cfg->stop_first = p->stop_on_first_hit; cfg->max_hits_to_get = p->max_hits_to_get;
    cfg->seed_coverage = p->seed_coverage;
    uint32_t ctor_seeds = p->num_seeds ? p->num_seeds : (uint32_t)(int)(p->seed_coverage * p->max_read_size / x->dev.seed_len);
    if (ctor_seeds == 0) return set_error(SNAPB200_ERR_ARG, "num_seeds/seed_coverage give zero seeds");
    cfg->n_lists = ctor_seeds + 1;
    cfg->rl = std::max(32u, (max_len + 15) & ~15u);
    return 0;
}

// Small tier: enough for almost every read; large tier: the bound implied by the reference's loop
// ((maxSeeds+1) seed directions of at most maxHits hits), with fewer resident warps if HBM would not hold it.
// Scratch tiers of the single-end aligner: every read first tries a small element pool at full occupancy; a read that
// overflows it is aborted and rerun from scratch in the next tier (deterministic, so the result cannot depend on the
// tier).  The last tier has the reference's own pool size (maxHits * (maxSeeds + 1) elements, BaseAligner.cpp:130), which
// at -h 16000 is 23 MB per warp and leaves room for only ~800 warps -- hence the middle tier, which still runs at full
// occupancy and takes all but a handful of the reads that outgrow the first.
#define MAX_SINGLE_TIERS 3
static int single_tiers(const snapb200_index *x, const SingleCfg &cfg, uint32_t max_len, SingleTier *tiers)
{
    uint32_t max_seeds = cfg.num_seeds ? cfg.num_seeds : (uint32_t)(int)(cfg.seed_coverage * max_len / x->dev.seed_len);
    uint64_t bound = (uint64_t)cfg.max_hits * (max_seeds + 1);
    if (bound < 64) bound = 64;
    if (bound > (1u << 24)) bound = 1u << 24;
    int per_sm = 1;
    size_t smem = single_warp_shared(cfg.rl) * WARPS_PER_CTA;
    int grid = grid_for(single_kernel, smem, x->sm_count, &per_sm);
    const uint64_t caps[MAX_SINGLE_TIERS] = {2048, 16384, bound};
    int n = 0;
    for (int i = 0; i < MAX_SINGLE_TIERS; i++) {
        const uint64_t cap = std::min<uint64_t>(caps[i], bound);
        if (n > 0 && cap <= tiers[n - 1].pool_cap) continue;
        SingleTier &t = tiers[n++];
        t.pool_cap = (uint32_t)cap;
        t.tsize = next_pow2((uint32_t)std::min<uint64_t>(cap * 2, 1u << 25));
        size_t per_warp = (size_t)t.pool_cap * sizeof(Elem) + (size_t)t.tsize * 2 * sizeof(int2);
        size_t warps = std::max<size_t>(1, SCRATCH_BUDGET / per_warp);
        t.grid = (int)std::min<size_t>((size_t)grid, std::max<size_t>(1, warps / WARPS_PER_CTA));
    }
    return n;
}
static int run_single_tiers(snapb200_session *s, const SingleCfg &cfg, uint32_t max_len, const DevBatch b[2], int two_batches,
                            const uint32_t *positions, uint32_t n_items, snapb200_single_result *results, int mapq_divisor)
{
    if (n_items == 0) return 0;
    SingleTier tiers[MAX_SINGLE_TIERS];
    const int n_tiers = single_tiers(s->idx, cfg, max_len, tiers);
    int rc;
    if ((rc = s->retry_list.ensure(((size_t)std::max(n_items, s->max_items) + 1) * 4))) return rc;
    DevBuf &tmp = s->retry_tmp;
    const uint32_t *pos = positions;
    uint32_t n = n_items;
    for (int t = 0; t < n_tiers; t++) {
        CUDA_TRY(cudaMemsetAsync((char *)s->counters.p + offsetof(Counters, n_retry), 0, 4, s->stream));
        if ((rc = launch_single(s, cfg, tiers[t], b, two_batches, pos, n, results, mapq_divisor))) break;
        if (t + 1 == n_tiers && n_tiers == 1) break;  // one tier = the reference's pool: nothing can overflow it
        Counters c;
        if ((rc = read_counters(s, &c))) break;
        if (c.n_retry == 0) break;
        if (t + 1 == n_tiers) { rc = set_error(SNAPB200_ERR_LIMIT, "%u reads overflowed the full-size candidate pool", c.n_retry); break; }
        // rerun the overflowed reads in the next tier; the retry list holds their result slots
        std::vector<uint32_t> list(c.n_retry);
        CUDA_TRY(cudaMemcpyAsync(list.data(), s->retry_list.p, (size_t)c.n_retry * 4, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        std::sort(list.begin(), list.end());
        if ((rc = tmp.ensure((size_t)c.n_retry * 4))) break;
        CUDA_TRY(cudaMemcpyAsync(tmp.p, list.data(), (size_t)c.n_retry * 4, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));  // `list` goes out of scope
        pos = tmp.as<uint32_t>();
        n = c.n_retry;
    }
    return rc;
}

static int begin_run(snapb200_session *s)
{
    CUDA_TRY(cudaSetDevice(s->idx->device));
    s->last_launches = 0;
    s->host_fix.clear();
    s->limit_hit = 0;
    s->main_pending = true;
    s->main_ms = 0;
    CUDA_TRY(cudaMemsetAsync(s->counters.p, 0, sizeof(Counters), s->stream));
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    return 0;
}

static int end_run(snapb200_session *s)
{
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    Counters c;
    int rc = read_counters(s, &c);
    if (rc) return rc;
    CUDA_TRY(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    if (!s->main_pending) CUDA_TRY(cudaEventElapsedTime(&s->main_ms, s->evm0, s->evm1));
    s->total_launches += s->last_launches;
    uint32_t nfix = std::min<uint32_t>(c.n_fix, FIX_CAP);
    if (c.n_fix > FIX_CAP) return set_error(SNAPB200_ERR_LIMIT, "too many mapq fix-up requests (%u)", c.n_fix);
    s->host_fix.resize(nfix);
    if (nfix) CUDA_TRY(cudaMemcpy(s->host_fix.data(), s->fix.p, sizeof(MapqFix) * nfix, cudaMemcpyDeviceToHost));
    s->limit_hit = c.n_limit != 0;
    return 0;
}

extern "C" int snapb200_session_run_single(snapb200_session *s, const snapb200_single_params *p)
{
    if (!s) return set_error(SNAPB200_ERR_ARG, "null session");
    SingleCfg cfg;
    int rc = single_cfg_from(s->idx, p, s->max_len_seen, &cfg);
    if (rc) return rc;
    if ((rc = begin_run(s))) return rc;
    const uint32_t n = s->n[0];
    if ((rc = s->single_res.ensure((size_t)std::max(n, 1u) * sizeof(snapb200_single_result)))) return rc;
    if (cfg.max_hits_to_get) {
        size_t mh = cfg.max_hits_to_get;
        if ((rc = s->mh_counts.ensure((size_t)std::max(n, 1u) * 4))) return rc;
        if ((rc = s->mh_locs.ensure((size_t)std::max(n, 1u) * mh * 4))) return rc;
        if ((rc = s->mh_rcs.ensure((size_t)std::max(n, 1u) * mh))) return rc;
        if ((rc = s->mh_scores.ensure((size_t)std::max(n, 1u) * mh * 4))) return rc;
    }
    DevBatch b[2] = {dev_batch(s, 0), dev_batch(s, 0)};
    const uint32_t *order = nullptr;  // heaviest reads first, as in the paired path
    if ((rc = work_order(s, 1, n, p->max_hits, &order))) return rc;
    if ((rc = run_single_tiers(s, cfg, s->max_len_seen, b, 0, order, n, s->single_res.as<snapb200_single_result>(), 1))) return rc;
    if (n) {
        stats_single_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(s->single_res.as<snapb200_single_result>(), n, s->idx->stats);
        s->last_launches++;
    }
    s->last_kind = 1;
    s->last_n = n;
    return end_run(s);
}

__global__ void expand_fallback_kernel(const uint32_t *fallback_list, uint32_t n, uint32_t *positions)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * n) positions[t] = fallback_list[t >> 1] * 2 + (t & 1);
}

static int launch_paired(snapb200_session *s, const snapb200_paired_params *p, const PairedCfg &cfg, int grid, const uint32_t *positions,
                         uint32_t n_items)
{
    snapb200_index *x = s->idx;
    PairedArgs a;
    memset(&a, 0, sizeof(a));
    a.ix_slot = x->slot;
    a.cfg = cfg;
    a.b[0] = dev_batch(s, 0);
    a.b[1] = dev_batch(s, 1);
    a.positions = positions;
    a.n_items = n_items;
    a.results = s->paired_res.as<snapb200_paired_result>();
    a.force_spacing = p->force_spacing;
    const size_t warps = (size_t)grid * WARPS_PER_CTA;
    int rc;
    if ((rc = s->p_cands.ensure(warps * cfg.cand_cap * sizeof(Cand)))) return rc;
    if ((rc = s->p_mates.ensure(warps * 2 * cfg.mate_cap * sizeof(Mate)))) return rc;
    if ((rc = s->p_anchors.ensure(warps * cfg.anchor_cap * sizeof(Anchor)))) return rc;
    if ((rc = s->p_lane_tables.ensure(warps * lane_table_cells((int)cfg.lane_k) * 32 * sizeof(lane_cell_t)))) return rc;
    if ((rc = s->p_order.ensure(warps * cfg.cand_cap * sizeof(uint32_t)))) return rc;
    a.cands = s->p_cands.as<Cand>(); a.mates = s->p_mates.as<Mate>(); a.anchors = s->p_anchors.as<Anchor>();
    a.lane_tables = s->p_lane_tables.as<lane_cell_t>();
    a.order = s->p_order.as<uint32_t>();
    a.ctr = s->counters.as<Counters>();
    a.retry_list = s->retry_list.as<uint32_t>(); a.fallback_list = s->fallback_list.as<uint32_t>();
    a.fix = s->fix.as<MapqFix>(); a.fix_cap = FIX_CAP;
    a.stats = x->stats;
    a.prof = nullptr;
    static DevBuf prof_buf;
#ifdef SNAPB200_PROFILE
    const bool prof = getenv("SNAPB200_PROF") != nullptr;
#else
    const bool prof = false;
#endif
    if (prof) {
        if ((rc = prof_buf.ensure(128))) return rc;
        CUDA_TRY(cudaMemsetAsync(prof_buf.p, 0, 128, s->stream));
        CUDA_TRY(cudaMemsetAsync((char *)prof_buf.p + 11 * 8, 0xff, 16, s->stream));  // the two minima
        a.prof = prof_buf.as<unsigned long long>();
    }
    if ((rc = reset_work(s))) return rc;
    const size_t smem = paired_warp_shared(cfg.rl, cfg.lane_k) * WARPS_PER_CTA;
    a.smem_per_warp = (uint32_t)paired_warp_shared(cfg.rl, cfg.lane_k);
    const bool time_it = s->main_pending;
    if (time_it) CUDA_TRY(cudaEventRecord(s->evm0, s->stream));
    paired_kernel<<<grid, CTA_THREADS, smem, s->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    if (time_it) { CUDA_TRY(cudaEventRecord(s->evm1, s->stream)); s->main_pending = false; }
    s->last_launches++;
    if (prof) {
        unsigned long long h[16];
        CUDA_TRY(cudaMemcpyAsync(h, prof_buf.p, 128, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        double tot = (double)h[0];
        {
            const double n_warps = (double)grid * WARPS_PER_CTA, span = (double)(h[13] - h[11]) * 1e-6, first_idle = (double)(h[12] - h[11]) * 1e-6;
            const double mean_exit = ((double)h[14] / n_warps - (double)h[11]) * 1e-6;
            fprintf(stderr, "[snapb200 prof]   timeline (globaltimer): first warp starts at 0, the first warp finds the queue empty at %.2f ms, the mean warp at %.2f ms, "
                            "the last at %.2f ms: warps are busy %.1f %% of the kernel\n", first_idle, mean_exit, span, 100 * mean_exit / span);
        }
        fprintf(stderr, "[snapb200 prof] paired_kernel grid=%d items=%llu cycles/item=%.0f  phase1 %.1f%%  phase2 %.1f%%  lv %.1f%%  leader3 %.1f%%\n", grid,
                h[5], h[5] ? tot / h[5] : 0.0, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot);
#ifdef SNAPB200_PROFILE
        {
            unsigned long long g[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
            cudaMemcpyFromSymbol(g, g_prof_lane, sizeof(g));
            cudaMemcpyToSymbol(g_prof_lane, z, sizeof(z));
            fprintf(stderr, "[snapb200 prof]   lane-mode calls %llu: live lanes forward %.1f, backward %.1f of 32\n", g[0], g[0] ? (double)g[1] / g[0] : 0.0, g[0] ? (double)g[2] / g[0] : 0.0);
        }
#endif
        fprintf(stderr, "[snapb200 prof]   lane mode: %llu batches, %.1f locations/batch, %.0f cycles/batch (%.1f%% of kernel cycles); warp mode: %llu calls, %.0f cycles/call (%.1f%%)\n",
                h[7], h[7] ? (double)h[8] / h[7] : 0.0, h[7] ? (double)h[6] / h[7] : 0.0, 100 * h[6] / tot, h[10], h[10] ? (double)h[9] / h[10] : 0.0, 100 * h[9] / tot);
    }
    return 0;
}

extern "C" int snapb200_session_run_paired(snapb200_session *s, const snapb200_paired_params *p)
{
    if (!s || !p) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (s->n[0] != s->n[1]) return set_error(SNAPB200_ERR_ARG, "mate batches differ in size (%u vs %u)", s->n[0], s->n[1]);
    snapb200_index *x = s->idx;
    if (p->max_k + p->extra_search_depth >= MAXK) return set_error(SNAPB200_ERR_ARG, "max_k + extra_search_depth must be < %d", MAXK);
    const uint32_t num_seeds = std::min(p->num_seeds, 30u);  // MAX_MAX_SEEDS (IntersectingPairedEndAligner.h:93)
    uint32_t ctor_seeds = num_seeds ? num_seeds : (uint32_t)(p->max_read_size * p->seed_coverage / x->dev.seed_len);
    uint32_t run_seeds = num_seeds ? num_seeds : (uint32_t)(s->max_len_seen * p->seed_coverage / x->dev.seed_len);
    if (run_seeds > MAX_LOOKUPS) return set_error(SNAPB200_ERR_ARG, "%u seeds per mate requested; this build holds %d lookups per hit set", run_seeds, MAX_LOOKUPS);
    if (ctor_seeds == 0) return set_error(SNAPB200_ERR_ARG, "num_seeds/seed_coverage give zero seeds");
    int rc;
    std::unique_lock<std::mutex> turn(x->run_turn, std::defer_lock);  // one session's kernels at a time on this index (see paired_chunks)
    if ((rc = begin_run(s))) return rc;
    const uint32_t n = s->n[0];
    const uint32_t rl = std::max(32u, (s->max_len_seen + 15) & ~15u);
    if ((rc = s->paired_res.ensure((size_t)std::max(n, 1u) * sizeof(snapb200_paired_result)))) return rc;
    if ((rc = s->retry_list.ensure(((size_t)std::max(2 * n, s->max_items) + 1) * 4))) return rc;
    if ((rc = s->fallback_list.ensure(((size_t)n + 1) * 4))) return rc;
    PairedCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.max_k = p->max_k; cfg.num_seeds = num_seeds; cfg.extra = p->extra_search_depth; cfg.min_spacing = p->min_spacing;
    cfg.max_spacing = p->max_spacing; cfg.max_big_hits = p->max_big_hits; cfg.seed_coverage = p->seed_coverage; cfg.rl = rl;
    // the reference's pool sizes (IntersectingPairedEndAligner.cpp:128-138)
    uint64_t ref_pool = std::min<uint64_t>(p->max_candidate_pool_size, (uint64_t)p->max_big_hits * ctor_seeds * 2);
    int per_sm = 1;
    cfg.lane_k = p->max_k + p->extra_search_depth;  // < MAXK (checked above)
    cfg.lane_gate = s->max_len_seen <= LANE_MAX_READ ? cfg.lane_k + 1 : 0;
    const size_t smem = paired_warp_shared(rl, cfg.lane_k) * WARPS_PER_CTA;
    int grid = grid_for(paired_kernel, smem, x->sm_count, &per_sm);
    if (const char *e = getenv("SNAPB200_CTAS_PER_SM")) {  // experiments only: fewer resident CTAs than fit
        int v = atoi(e);
        if (v >= 1 && v < per_sm) { per_sm = v; grid = v * x->sm_count; }
    }
    if (n) {
        // small tier
        // first tier: ~1.4 MB of candidate records per warp (5 GB at full occupancy) holds all but pathological pairs
        cfg.cand_cap = (uint32_t)std::min<uint64_t>(ref_pool, 12288);
        cfg.mate_cap = (uint32_t)std::min<uint64_t>(ref_pool / 2, 12288);
        cfg.anchor_cap = cfg.cand_cap;
        cfg.hard_limit = (cfg.cand_cap == ref_pool && cfg.mate_cap == ref_pool / 2) ? 1 : 0;
        // heaviest pairs first (see weigh_kernel); SNAPB200_NO_ORDER=1 serves them in input order (experiments)
        const uint32_t *order = nullptr;
        if ((rc = work_order(s, 2, n, p->max_big_hits, &order))) return rc;
        // the ordering kernels are queued before the turn is taken: they run as soon as the other session's main kernel leaves room,
        // next to its short post-main kernels, instead of after them
        if (turns_enabled()) turn.lock();
        if ((rc = launch_paired(s, p, cfg, grid, order, n))) return rc;
        Counters c;
        if ((rc = read_counters(s, &c))) return rc;
        if (c.n_retry) {
            std::vector<uint32_t> list(c.n_retry);
            CUDA_TRY(cudaMemcpyAsync(list.data(), s->retry_list.p, (size_t)c.n_retry * 4, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            std::sort(list.begin(), list.end());
            DevBuf &tmp = s->retry_tmp;
            if ((rc = tmp.ensure((size_t)c.n_retry * 4))) return rc;
            CUDA_TRY(cudaMemcpyAsync(tmp.p, list.data(), (size_t)c.n_retry * 4, cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaMemsetAsync((char *)s->counters.p + offsetof(Counters, n_retry), 0, 4, s->stream));
            PairedCfg big = cfg;
            big.cand_cap = (uint32_t)ref_pool; big.mate_cap = (uint32_t)(ref_pool / 2); big.anchor_cap = (uint32_t)ref_pool;
            big.hard_limit = 1;
            size_t per_warp = (size_t)big.cand_cap * (sizeof(Cand) + sizeof(uint32_t)) + (size_t)big.mate_cap * 2 * sizeof(Mate) + (size_t)big.anchor_cap * sizeof(Anchor);
            size_t warps = std::max<size_t>(1, SCRATCH_BUDGET / std::max<size_t>(per_warp, 1));
            int g = (int)std::min<size_t>((size_t)grid, std::max<size_t>(1, warps / WARPS_PER_CTA));
            rc = launch_paired(s, p, big, g, tmp.as<uint32_t>(), c.n_retry);
            if (!rc) rc = read_counters(s, &c);
            if (rc) return rc;
        }
        // single-end fallback for the pairs the intersecting aligner could not place
        if (c.n_fallback) {
            const uint32_t nf = c.n_fallback;
            if ((rc = s->fb_single_res.ensure((size_t)2 * n * sizeof(snapb200_single_result)))) return rc;
            if ((rc = s->fb_positions.ensure((size_t)2 * nf * 4))) return rc;
            expand_fallback_kernel<<<(2 * nf + 255) / 256, 256, 0, s->stream>>>(s->fallback_list.as<uint32_t>(), nf, s->fb_positions.as<uint32_t>());
            s->last_launches++;
            snapb200_single_params sp;
            memset(&sp, 0, sizeof(sp));
            sp.max_hits = p->max_hits; sp.max_k = p->max_k; sp.max_read_size = p->max_read_size; sp.num_seeds = p->num_seeds;
            sp.seed_coverage = p->seed_coverage; sp.extra_search_depth = p->extra_search_depth;
            SingleCfg scfg;
            if ((rc = single_cfg_from(x, &sp, s->max_len_seen, &scfg))) return rc;
            DevBatch b[2] = {dev_batch(s, 0), dev_batch(s, 1)};
            if ((rc = run_single_tiers(s, scfg, s->max_len_seen, b, 1, s->fb_positions.as<uint32_t>(), 2 * nf,
                                       s->fb_single_res.as<snapb200_single_result>(), 4))) return rc;
            merge_fallback_kernel<<<(2 * nf + 255) / 256, 256, 0, s->stream>>>(s->fallback_list.as<uint32_t>(), nf,
                                                                               s->fb_single_res.as<snapb200_single_result>(),
                                                                               s->paired_res.as<snapb200_paired_result>());
            s->last_launches++;
        }
        stats_paired_kernel<<<(2 * n + 255) / 256, 256, 0, s->stream>>>(s->paired_res.as<snapb200_paired_result>(), n, x->stats);
        s->last_launches++;
    }
    s->last_kind = 2;
    s->last_n = n;
    return end_run(s);
}

extern "C" int snapb200_session_download_single(snapb200_session *s, snapb200_single_result *results)
{
    if (!s || s->last_kind != 1) return set_error(SNAPB200_ERR_ARG, "no single-end run to download");
    CUDA_TRY(cudaSetDevice(s->idx->device));
    if (s->last_n) {
        CUDA_TRY(cudaMemcpyAsync(results, s->single_res.p, (size_t)s->last_n * sizeof(snapb200_single_result), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    for (const MapqFix &f : s->host_fix) {  // see compute_mapq_dev
        snapb200_single_result *r = &results[f.index];
        int mq = compute_mapq_host(f.p_all, f.p_best, f.score, f.popular);
        r->mapq = mq;
        r->status = mq >= 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
    }
    return 0;
}

extern "C" int snapb200_session_download_paired(snapb200_session *s, snapb200_paired_result *results)
{
    if (!s || s->last_kind != 2) return set_error(SNAPB200_ERR_ARG, "no paired-end run to download");
    CUDA_TRY(cudaSetDevice(s->idx->device));
    if (s->last_n) {
        CUDA_TRY(cudaMemcpyAsync(results, s->paired_res.p, (size_t)s->last_n * sizeof(snapb200_paired_result), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    for (const MapqFix &f : s->host_fix) {
        int mq = compute_mapq_host(f.p_all, f.p_best, f.score, f.popular);
        if (f.is_paired_rule) {
            snapb200_paired_result *r = &results[f.index];
            r->mapq[f.end] = mq;
            r->status[f.end] = mq > 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
        } else {  // single-end fallback: index = pair*2+end, mapq/4 (ChimericPairedEndAligner.cpp:115)
            snapb200_paired_result *r = &results[f.index >> 1];
            r->status[f.index & 1] = mq >= 10 ? SNAPB200_SINGLE_HIT : SNAPB200_MULTIPLE_HITS;
            r->mapq[f.index & 1] = mq / f.divisor;
        }
    }
    if (s->limit_hit) return set_error(SNAPB200_ERR_LIMIT, "a pair exceeded the reference's candidate pool (-mcp); the reference exits here");
    return 0;
}

extern "C" int snapb200_session_sync(snapb200_session *s)
{
    if (!s) return set_error(SNAPB200_ERR_ARG, "null session");
    CUDA_TRY(cudaSetDevice(s->idx->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int snapb200_session_last_run(const snapb200_session *s, float *kernel_ms, uint32_t *launches, uint64_t *total_launches)
{
    if (!s) return set_error(SNAPB200_ERR_ARG, "null session");
    if (kernel_ms) *kernel_ms = s->last_ms;
    if (launches) *launches = s->last_launches;
    if (total_launches) *total_launches = s->total_launches;
    return 0;
}

extern "C" int snapb200_session_main_kernel_ms(const snapb200_session *s, float *main_ms)
{
    if (!s || !main_ms) return set_error(SNAPB200_ERR_ARG, "null argument");
    *main_ms = s->main_ms;
    return 0;
}

// ---- synchronous batch entry points: chunked, double-buffered over two sessions ------------------------------
static const uint32_t CHUNK_DEFAULT = 1u << 20;
// reads (pairs) per internal launch group; SNAPB200_CHUNK overrides it (used by the tests to force multi-chunk batches)
static uint32_t chunk_size()
{
    const char *e = getenv("SNAPB200_CHUNK");
    if (e) {
        long v = atol(e);
        if (v >= 1024 && v <= (1l << 24)) return (uint32_t)v;
    }
    return CHUNK_DEFAULT;
}

static snapb200_read_batch sub_batch(const snapb200_read_batch *b, uint32_t lo, uint32_t hi, std::vector<uint32_t> &off_store)
{
    snapb200_read_batch r;
    r.n = hi - lo;
    const uint32_t base = b->offsets[lo];
    if (base == 0) {
        r.offsets = b->offsets + lo;
    } else {
        off_store.resize(r.n + 1);
        for (uint32_t i = 0; i <= r.n; i++) off_store[i] = b->offsets[lo + i] - base;
        r.offsets = off_store.data();
    }
    r.bases = b->bases + base;
    r.quals = b->quals + base;
    return r;
}

// RAII: one of the index's two internal sessions, locked for the duration of a *_batch call
struct BatchSlot {
    snapb200_index *idx;
    int slot = -1;
    snapb200_session *s = nullptr;
    explicit BatchSlot(snapb200_index *i) : idx(i)
    {
        const int n = batch_sessions();
        for (int k = 0; k < n && slot < 0; k++) if (idx->batch_mutex[k].try_lock()) slot = k;
        if (slot < 0) {
            slot = (int)(idx->batch_rr.fetch_add(1) % (unsigned)n);
            idx->batch_mutex[slot].lock();
        }
    }
    int open()
    {
        if (!idx->batch_session[slot]) {
            int rc = snapb200_session_create(idx, CHUNK_DEFAULT, SNAPB200_MAX_READ_LENGTH, &idx->batch_session[slot]);
            if (rc) return rc;
        }
        s = idx->batch_session[slot];
        return 0;
    }
    ~BatchSlot() { idx->batch_mutex[slot].unlock(); }
};

static int single_batch_impl(snapb200_index *idx, const snapb200_single_params *params, const snapb200_read_batch *reads,
                             snapb200_single_result *results, int32_t *hit_counts, uint32_t *hit_locations, uint8_t *hit_rcs,
                             int32_t *hit_scores)
{
    if (!idx || !params || !reads || (reads->n && !results)) return set_error(SNAPB200_ERR_ARG, "null argument");
    uint32_t m;
    int rc = validate_batch(reads, &m);
    if (rc) return rc;
    BatchSlot slot(idx);
    if ((rc = slot.open())) return rc;
    snapb200_session *cur = slot.s;
    const uint32_t n = reads->n;
    const uint32_t mh = params->max_hits_to_get;
    std::vector<uint32_t> off_store;
    const uint32_t CHUNK = chunk_size();
    for (uint32_t lo = 0; lo < n; lo += CHUNK) {
        const uint32_t hi = std::min(n, lo + CHUNK);
        snapb200_read_batch sb = sub_batch(reads, lo, hi, off_store);
        if ((rc = snapb200_session_upload(cur, 0, &sb))) return rc;
        if ((rc = snapb200_session_run_single(cur, params))) return rc;
        if ((rc = snapb200_session_download_single(cur, results + lo))) return rc;
        if (mh) {
            const uint32_t k = hi - lo;
            CUDA_TRY(cudaMemcpy(hit_counts + lo, cur->mh_counts.p, (size_t)k * 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(hit_locations + (size_t)lo * mh, cur->mh_locs.p, (size_t)k * mh * 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(hit_rcs + (size_t)lo * mh, cur->mh_rcs.p, (size_t)k * mh, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(hit_scores + (size_t)lo * mh, cur->mh_scores.p, (size_t)k * mh * 4, cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

extern "C" int snapb200_single_batch(snapb200_index *idx, const snapb200_single_params *params, const snapb200_read_batch *reads,
                                     snapb200_single_result *results)
{
    if (!params) return set_error(SNAPB200_ERR_ARG, "null params");
    snapb200_single_params p = *params;
    p.max_hits_to_get = 0;
    return single_batch_impl(idx, &p, reads, results, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int snapb200_single_multihit_batch(snapb200_index *idx, const snapb200_single_params *params,
                                              const snapb200_read_batch *reads, snapb200_single_result *results,
                                              int32_t *hit_counts, uint32_t *hit_locations, uint8_t *hit_rcs, int32_t *hit_scores)
{
    if (!params) return set_error(SNAPB200_ERR_ARG, "null params");
    if (params->max_hits_to_get && (!hit_counts || !hit_locations || !hit_rcs || !hit_scores)) return set_error(SNAPB200_ERR_ARG, "null multi-hit output");
    return single_batch_impl(idx, params, reads, results, hit_counts, hit_locations, hit_rcs, hit_scores);
}

// One worker of snapb200_paired_batch: chunks first, first + stride, ... on one of the index's internal sessions.
static int paired_chunks(snapb200_index *idx, const snapb200_paired_params *params, const snapb200_read_batch *reads0, const snapb200_read_batch *reads1,
                         snapb200_paired_result *results, uint32_t first, uint32_t stride, uint32_t CHUNK, int *limit_rc)
{
    BatchSlot slot(idx);
    int rc;
    if ((rc = slot.open())) return rc;
    snapb200_session *cur = slot.s;
    const uint32_t n = reads0->n;
    std::vector<uint32_t> off_store[2];
    for (uint64_t lo64 = (uint64_t)first * CHUNK; lo64 < n; lo64 += (uint64_t)stride * CHUNK) {
        const uint32_t lo = (uint32_t)lo64, hi = (uint32_t)std::min<uint64_t>(n, lo64 + CHUNK);
        snapb200_read_batch a = sub_batch(reads0, lo, hi, off_store[0]);
        snapb200_read_batch b = sub_batch(reads1, lo, hi, off_store[1]);
        static const bool timing = getenv("SNAPB200_BATCH_TIMING") != nullptr;  // where a chunk's wall time goes (stderr)
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = snapb200_session_upload(cur, 0, &a)) || (rc = snapb200_session_upload(cur, 1, &b))) return rc;
        const auto t1 = std::chrono::steady_clock::now();
        // The two sessions take turns on the SMs.  The persistent main kernel fills the device, so the short kernels that follow a
        // main kernel (scratch-tier retries, the single-end fallback) could not start before the OTHER session's main kernel had
        // finished: both sessions then completed together, uploaded together, and the device sat idle for the length of an upload in
        // every cycle (device timeline in profiles/README.md).  With the turn, a chunk's kernels run back to back and the other
        // session's download and upload fall entirely under them.  The upload is waited for first, so a turn never starts with a copy.
        // (the turn itself is taken inside snapb200_session_run_paired, so that sessions driven directly take turns as well)
        {
            cudaError_t e = turns_enabled() ? cudaStreamSynchronize(cur->stream) : cudaSuccess;
            if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
            if ((rc = snapb200_session_run_paired(cur, params))) return rc;
        }
        const auto t2 = std::chrono::steady_clock::now();
        rc = snapb200_session_download_paired(cur, results + lo);
        if (timing) {
            const auto t3 = std::chrono::steady_clock::now();
            auto ms = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
            static cudaEvent_t base = nullptr;  // device-side timeline: when this chunk's first and main kernels ran, on one clock for all sessions
            static std::mutex base_lock;
            {
                std::lock_guard<std::mutex> g(base_lock);
                if (!base) { cudaEventCreate(&base); cudaEventRecord(base, cur->stream); cudaEventSynchronize(base); }
            }
            float e0 = 0, m0 = 0, m1 = 0, e1 = 0;
            cudaEventElapsedTime(&e0, base, cur->ev0); cudaEventElapsedTime(&m0, base, cur->evm0);
            cudaEventElapsedTime(&m1, base, cur->evm1); cudaEventElapsedTime(&e1, base, cur->ev1);
            fprintf(stderr, "[snapb200 batch] %u pairs on session %d: upload %.1f ms, run %.1f ms (kernels of this chunk %.1f ms, main kernel %.1f ms), download %.1f ms; "
                            "device timeline: inputs resident at %.1f, main kernel %.1f .. %.1f, last kernel done %.1f\n",
                    hi - lo, slot.slot, ms(t0, t1), ms(t1, t2), cur->last_ms, cur->main_ms, ms(t2, t3), e0, m0, m1, e1);
        }
        if (rc == SNAPB200_ERR_LIMIT) *limit_rc = rc; else if (rc) return rc;
    }
    return 0;
}

extern "C" int snapb200_paired_batch(snapb200_index *idx, const snapb200_paired_params *params, const snapb200_read_batch *reads0,
                                     const snapb200_read_batch *reads1, snapb200_paired_result *results)
{
    if (!idx || !params || !reads0 || !reads1 || (reads0->n && !results)) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (reads0->n != reads1->n) return set_error(SNAPB200_ERR_ARG, "mate batches differ in size");
    uint32_t m0, m1;
    int rc;
    if ((rc = validate_batch(reads0, &m0)) || (rc = validate_batch(reads1, &m1))) return rc;
    const uint32_t CHUNK = chunk_size();
    int limit_rc[2] = {0, 0};
    if (reads0->n <= CHUNK) {
        rc = paired_chunks(idx, params, reads0, reads1, results, 0, 1, CHUNK, &limit_rc[0]);
        return rc ? rc : limit_rc[0];
    }
    // More than one launch group: even and odd chunks go through the two internal sessions from two host threads, so the upload
    // of chunk k+1 and the download of chunk k-1 (and the host round trips inside a run: scratch-tier retries, counters) overlap
    // the kernels of chunk k -- what two concurrent callers get, inside one call.
    int rc2 = 0;
    char err2[sizeof(g_last_error)] = "";
    std::thread odd([&] {
        rc2 = paired_chunks(idx, params, reads0, reads1, results, 1, 2, CHUNK, &limit_rc[1]);
        if (rc2) { strncpy(err2, g_last_error, sizeof(err2) - 1); err2[sizeof(err2) - 1] = 0; }
    });
    rc = paired_chunks(idx, params, reads0, reads1, results, 0, 2, CHUNK, &limit_rc[0]);
    odd.join();
    if (!rc && rc2) { rc = rc2; strncpy(g_last_error, err2, sizeof(g_last_error) - 1); }
    if (rc) return rc;
    if (limit_rc[0] || limit_rc[1]) return set_error(SNAPB200_ERR_LIMIT, "a pair exceeded the reference's candidate pool (-mcp); the reference exits here");
    return 0;
}

// ---- CIGAR -----------------------------------------------------------------------------------------------------
extern "C" int snapb200_cigar_batch(snapb200_index *idx, const snapb200_read_batch *reads, const uint32_t *locations,
                                    const uint8_t *directions, int use_m, char *cigars, uint32_t cigar_stride, int32_t *edit_distance)
{
    if (!idx || !reads || (reads->n && (!locations || !directions || !cigars || !edit_distance))) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (cigar_stride < 2) return set_error(SNAPB200_ERR_ARG, "cigar_stride too small");
    uint32_t m;
    int rc = validate_batch(reads, &m);
    if (rc) return rc;
    const uint32_t n = reads->n;
    if (!n) return 0;
    BatchSlot slot(idx);
    if ((rc = slot.open())) return rc;
    snapb200_session *ss = slot.s;
    if ((rc = snapb200_session_upload(ss, 0, reads))) return rc;
    DevBuf d_loc, d_dir, d_cig, d_ed;
    do {
        if ((rc = d_loc.ensure((size_t)n * 4)) || (rc = d_dir.ensure(n)) || (rc = d_cig.ensure((size_t)n * cigar_stride)) || (rc = d_ed.ensure((size_t)n * 4))) break;
        cudaMemcpyAsync(d_loc.p, locations, (size_t)n * 4, cudaMemcpyHostToDevice, ss->stream);
        cudaMemcpyAsync(d_dir.p, directions, n, cudaMemcpyHostToDevice, ss->stream);
        CigarArgs a;
        memset(&a, 0, sizeof(a));
        a.ix = idx->dev; a.b = dev_batch(ss, 0); a.locations = d_loc.as<uint32_t>(); a.directions = d_dir.as<uint8_t>();
        a.use_m = use_m; a.cigars = d_cig.as<char>(); a.stride = cigar_stride; a.edit_distance = d_ed.as<int32_t>();
        a.rl = std::max(32u, (m + 15) & ~15u); a.ctr = ss->counters.as<Counters>();
        if ((rc = reset_work(ss))) break;
        size_t smem = cigar_warp_shared(a.rl) * WARPS_PER_CTA;
        int per_sm;
        int grid = grid_for(cigar_kernel, smem, idx->sm_count, &per_sm);
        cigar_kernel<<<grid, CTA_THREADS, smem, ss->stream>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "cigar_kernel launch: %s", cudaGetErrorString(e)); break; }
        ss->total_launches++;
        cudaMemcpyAsync(cigars, d_cig.p, (size_t)n * cigar_stride, cudaMemcpyDeviceToHost, ss->stream);
        cudaMemcpyAsync(edit_distance, d_ed.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ss->stream);
        e = cudaStreamSynchronize(ss->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "cigar_kernel: %s", cudaGetErrorString(e)); break; }
    } while (0);
    d_loc.release(); d_dir.release(); d_cig.release(); d_ed.release();
    return rc;
}

// ---- CharacterizeSeeds -------------------------------------------------------------------------------------------
// One launch group over the batch resident in slot `mate_slot` of `ss`; the results stay in HBM: d_off = exclusive segment offsets
// [2k + 1] (the last one is the number of tuples, also returned), d_loc / d_so = the ordered tuples.  The stream is idle on return
// only as far as the tuple count needed it; the caller synchronises before reading the buffers from the host.
static int characterize_device(snapb200_index *idx, snapb200_session *ss, int mate_slot, const snapb200_single_params *params, uint32_t max_len,
                               DevBuf &d_cnt, DevBuf &d_off, DevBuf &d_keys0, DevBuf &d_keys1, DevBuf &d_tmp, DevBuf &d_loc, DevBuf &d_so, uint64_t *total_out,
                               bool count_only = false)
{
    const uint32_t k = ss->n[mate_slot];
    const size_t n_seg = (size_t)2 * k;
    int rc;
    if ((rc = d_cnt.ensure((n_seg + 1) * 8)) || (rc = d_off.ensure((n_seg + 1) * 8))) return rc;
    CharArgs a;
    memset(&a, 0, sizeof(a));
    a.ix = idx->dev; a.b = dev_batch(ss, mate_slot);
    a.max_hits = params->max_hits; a.max_k = params->max_k; a.num_seeds = params->num_seeds;
    a.explore = params->explore_popular_seeds; a.seed_coverage = params->seed_coverage;
    a.rl = std::max(32u, (max_len + 15) & ~15u);
    a.ctr = ss->counters.as<Counters>();
    const size_t smem = char_warp_shared(a.rl) * WARPS_PER_CTA;
    int per_sm;
    const int grid_c = grid_for(characterize_kernel<false>, smem, idx->sm_count, &per_sm);
    const int grid_e = grid_for(characterize_kernel<true>, smem, idx->sm_count, &per_sm);
    // pass 1: segment sizes
    CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, (n_seg + 1) * 8, ss->stream));
    if ((rc = reset_work(ss))) return rc;
    a.seg = d_cnt.as<unsigned long long>();
    characterize_kernel<false><<<grid_c, CTA_THREADS, smem, ss->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "characterize_kernel launch: %s", cudaGetErrorString(e));
    ss->total_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt.as<unsigned long long>(), d_off.as<unsigned long long>(), n_seg + 1, ss->stream);
    if ((rc = d_tmp.ensure(tmp_bytes))) return rc;
    e = cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_cnt.as<unsigned long long>(), d_off.as<unsigned long long>(), n_seg + 1, ss->stream);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "characterize scan: %s", cudaGetErrorString(e));
    unsigned long long total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_off.as<unsigned long long>() + n_seg, 8, cudaMemcpyDeviceToHost, ss->stream));
    e = cudaStreamSynchronize(ss->stream);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "characterize_kernel: %s", cudaGetErrorString(e));
    *total_out = total;
    if (!total || count_only) return 0;
    if ((rc = d_keys0.ensure(total * 8)) || (rc = d_keys1.ensure(total * 8)) || (rc = d_loc.ensure(total * 4)) || (rc = d_so.ensure(total * 2))) return rc;
    // pass 2: one key per tuple, then order every segment
    if ((rc = reset_work(ss))) return rc;
    a.seg = d_off.as<unsigned long long>();
    a.keys = d_keys0.as<unsigned long long>();
    characterize_kernel<true><<<grid_e, CTA_THREADS, smem, ss->stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "characterize_kernel launch: %s", cudaGetErrorString(e));
    ss->total_launches++;
    int seg_bits = 1;
    while (((size_t)1 << seg_bits) < n_seg) seg_bits++;
    cub::DoubleBuffer<unsigned long long> kb(d_keys0.as<unsigned long long>(), d_keys1.as<unsigned long long>());
    tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, kb, (unsigned long long)total, 0, CHAR_SEG_SHIFT + seg_bits, ss->stream);
    if ((rc = d_tmp.ensure(tmp_bytes))) return rc;
    e = cub::DeviceRadixSort::SortKeys(d_tmp.p, tmp_bytes, kb, (unsigned long long)total, 0, CHAR_SEG_SHIFT + seg_bits, ss->stream);
    if (e != cudaSuccess) return set_error(SNAPB200_ERR_CUDA, "characterize sort: %s", cudaGetErrorString(e));
    const unsigned blocks = (unsigned)((total + 255) / 256);
    characterize_split_kernel<<<blocks, 256, 0, ss->stream>>>(kb.Current(), total, d_loc.as<uint32_t>(), d_so.as<uint16_t>());
    ss->total_launches++;
    return 0;
}

extern "C" int snapb200_characterize_batch(snapb200_index *idx, const snapb200_single_params *params, const snapb200_read_batch *reads,
                                           uint64_t *seg_offsets, uint32_t *locations, uint16_t *seed_offsets, uint64_t capacity)
{
    if (!idx || !params || !reads || !seg_offsets) return set_error(SNAPB200_ERR_ARG, "null argument");
    if ((locations == nullptr) != (seed_offsets == nullptr)) return set_error(SNAPB200_ERR_ARG, "locations and seed_offsets must both be given or both be NULL");
    uint32_t m;
    int rc = validate_batch(reads, &m);
    if (rc) return rc;
    if (params->max_k > MAXK) return set_error(SNAPB200_ERR_ARG, "max_k %u > MAX_K", params->max_k);
    const uint32_t n = reads->n;
    seg_offsets[0] = 0;
    if (!n) return 0;
    BatchSlot slot(idx);
    if ((rc = slot.open())) return rc;
    snapb200_session *ss = slot.s;
    CUDA_TRY(cudaSetDevice(idx->device));
    // reads per launch group: segment ids must fit the key bits above location and seed offset
    const uint32_t CHUNK = std::min(chunk_size(), 1u << (64 - CHAR_SEG_SHIFT - 2));
    std::vector<uint32_t> off_store;
    DevBuf d_seg, d_off, d_keys0, d_keys1, d_tmp, d_loc, d_so;
    uint64_t base = 0;
    for (uint32_t lo = 0; lo < n && !rc; lo += CHUNK) {
        const uint32_t hi = std::min(n, lo + CHUNK), k = hi - lo;
        snapb200_read_batch sb = sub_batch(reads, lo, hi, off_store);
        if ((rc = snapb200_session_upload(ss, 0, &sb))) break;
        const size_t n_seg = (size_t)2 * k;
        uint64_t total = 0;
        if ((rc = characterize_device(idx, ss, 0, params, m, d_seg, d_off, d_keys0, d_keys1, d_tmp, d_loc, d_so, &total, locations == nullptr))) break;
        cudaMemcpyAsync(seg_offsets + (size_t)2 * lo, d_off.p, (n_seg + 1) * 8, cudaMemcpyDeviceToHost, ss->stream);
        if (locations && total) {
            if (base + total > capacity) {
                cudaStreamSynchronize(ss->stream);
                if (base) for (size_t q = 0; q <= n_seg; q++) seg_offsets[(size_t)2 * lo + q] += base;
                rc = set_error(SNAPB200_ERR_ARG, "characterize: output capacity %llu too small (reads %u..%u alone need %llu more tuples)",
                               (unsigned long long)capacity, lo, hi, (unsigned long long)total);
                break;
            }
            cudaMemcpyAsync(locations + base, d_loc.p, total * 4, cudaMemcpyDeviceToHost, ss->stream);
            cudaMemcpyAsync(seed_offsets + base, d_so.p, total * 2, cudaMemcpyDeviceToHost, ss->stream);
        }
        cudaError_t e = cudaStreamSynchronize(ss->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "characterize emit: %s", cudaGetErrorString(e)); break; }
        if (base) for (size_t q = 0; q <= n_seg; q++) seg_offsets[(size_t)2 * lo + q] += base;
        base += total;
    }
    d_seg.release(); d_off.release(); d_keys0.release(); d_keys1.release(); d_tmp.release(); d_loc.release(); d_so.release();
    return rc;
}

// ---- building blocks --------------------------------------------------------------------------------------------
struct TableOnlyIndex {  // probability tables on a device, for the explicit-string LV entry points
    int device = -1;
    snapb200_index *x = nullptr;
};
static TableOnlyIndex g_tables[16];

static int tables_for(int device, snapb200_index **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(SNAPB200_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev || device >= 16) return set_error(SNAPB200_ERR_ARG, "device %d out of range", device);
    if (!g_tables[device].x) {
        CUDA_TRY(cudaSetDevice(device));
        snapb200_index *x = new snapb200_index();
        x->device = device;
        memset(&x->dev, 0, sizeof(x->dev));
        memset(&x->info, 0, sizeof(x->info));
        x->dev.seed_len = 20;
        int rc = finish_index(x);
        if (rc) { snapb200_index_close(x); return rc; }
        g_tables[device].x = x;
    }
    *out = g_tables[device].x;
    CUDA_TRY(cudaSetDevice(device));
    return 0;
}

static int lv_common(int device, int dir, uint32_t n, const uint32_t *text_offsets, const uint8_t *texts, const uint32_t *pattern_offsets,
                     const uint8_t *patterns, const uint8_t *quals, const int32_t *k, int32_t *score, double *prob, int32_t *indel,
                     int use_m, char *cigars, uint32_t stride)
{
    snapb200_index *x;
    int rc = tables_for(device, &x);
    if (rc) return rc;
    if (!n) return 0;
    if (!text_offsets || !pattern_offsets || !texts || !patterns || !k || !score) return set_error(SNAPB200_ERR_ARG, "null argument");
    uint32_t max_t = 0, max_p = 0;
    for (uint32_t i = 0; i < n; i++) {
        max_t = std::max(max_t, text_offsets[i + 1] - text_offsets[i]);
        max_p = std::max(max_p, pattern_offsets[i + 1] - pattern_offsets[i]);
    }
    if (max_p > SNAPB200_MAX_READ_LENGTH || max_t > SNAPB200_MAX_READ_LENGTH + 64) return set_error(SNAPB200_ERR_ARG, "string too long for the LV test entry point");
    const size_t tb = text_offsets[n], pb = pattern_offsets[n];
    DevBuf d_to, d_po, d_t, d_p, d_q, d_k, d_s, d_pr, d_in, d_c, d_ctr;
    cudaStream_t st = x->stream;
    do {
        if ((rc = d_to.ensure((size_t)(n + 1) * 4)) || (rc = d_po.ensure((size_t)(n + 1) * 4)) || (rc = d_t.ensure(tb + 16)) || (rc = d_p.ensure(pb + 16)) ||
            (rc = d_q.ensure(pb + 16)) || (rc = d_k.ensure((size_t)n * 4)) || (rc = d_s.ensure((size_t)n * 4)) || (rc = d_pr.ensure((size_t)n * 8)) ||
            (rc = d_in.ensure((size_t)n * 4)) || (rc = d_ctr.ensure(sizeof(Counters)))) break;
        if (cigars && (rc = d_c.ensure((size_t)n * stride))) break;
        cudaMemcpyAsync(d_to.p, text_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_po.p, pattern_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st);
        if (tb) cudaMemcpyAsync(d_t.p, texts, tb, cudaMemcpyHostToDevice, st);
        if (pb) cudaMemcpyAsync(d_p.p, patterns, pb, cudaMemcpyHostToDevice, st);
        if (pb && quals) cudaMemcpyAsync(d_q.p, quals, pb, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_k.p, k, (size_t)n * 4, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(d_ctr.p, 0, sizeof(Counters), st);
        LvArgs a;
        memset(&a, 0, sizeof(a));
        a.ix_slot = x->slot; a.dir = dir; a.n = n; a.text_off = d_to.as<uint32_t>(); a.pat_off = d_po.as<uint32_t>();
        a.texts = d_t.as<uint8_t>(); a.pats = d_p.as<uint8_t>(); a.quals = quals ? d_q.as<uint8_t>() : nullptr; a.k = d_k.as<int32_t>();
        a.score = d_s.as<int32_t>(); a.indel = d_in.as<int32_t>(); a.prob = d_pr.as<double>();
        a.use_m = use_m; a.cigars = cigars ? d_c.as<char>() : nullptr; a.stride = stride;
        a.max_text = std::max(16u, max_t); a.max_pat = std::max(16u, max_p); a.ctr = d_ctr.as<Counters>();
        size_t smem = lvtest_warp_shared(a.max_text, a.max_pat) * WARPS_PER_CTA;
        int per_sm;
        int grid = grid_for(lv_kernel, smem, x->sm_count, &per_sm);
        lv_kernel<<<grid, CTA_THREADS, smem, st>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "lv_kernel launch: %s", cudaGetErrorString(e)); break; }
        cudaMemcpyAsync(score, d_s.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
        if (prob) cudaMemcpyAsync(prob, d_pr.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
        if (indel) cudaMemcpyAsync(indel, d_in.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
        if (cigars) cudaMemcpyAsync(cigars, d_c.p, (size_t)n * stride, cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "lv_kernel: %s", cudaGetErrorString(e)); break; }
    } while (0);
    DevBuf *all[] = {&d_to, &d_po, &d_t, &d_p, &d_q, &d_k, &d_s, &d_pr, &d_in, &d_c, &d_ctr};
    for (DevBuf *b : all) b->release();
    return rc;
}

extern "C" int snapb200_lv_batch(int device, int text_direction, uint32_t n, const uint32_t *text_offsets, const uint8_t *texts,
                                 const uint32_t *pattern_offsets, const uint8_t *patterns, const uint8_t *quals, const int32_t *k,
                                 int32_t *score, double *match_probability, int32_t *net_indel)
{
    if (text_direction != 1 && text_direction != -1) return set_error(SNAPB200_ERR_ARG, "text_direction must be +1 or -1");
    return lv_common(device, text_direction, n, text_offsets, texts, pattern_offsets, patterns, quals, k, score, match_probability, net_indel, 0,
                     nullptr, 0);
}

extern "C" int snapb200_lv_cigar_batch(int device, uint32_t n, const uint32_t *text_offsets, const uint8_t *texts,
                                       const uint32_t *pattern_offsets, const uint8_t *patterns, const int32_t *k, int use_m, char *cigars,
                                       uint32_t cigar_stride, int32_t *edit_distance)
{
    if (n && (!cigars || cigar_stride < 2)) return set_error(SNAPB200_ERR_ARG, "bad cigar buffer");
    return lv_common(device, 1, n, text_offsets, texts, pattern_offsets, patterns, nullptr, k, edit_distance, nullptr, nullptr, use_m, cigars,
                     cigar_stride);
}

extern "C" int snapb200_lookup_seed_batch(snapb200_index *idx, uint32_t n, const uint8_t *seeds, uint32_t max_out, uint32_t *n_hits,
                                          uint32_t *hits)
{
    if (!idx || (n && (!seeds || !n_hits || (max_out && !hits)))) return set_error(SNAPB200_ERR_ARG, "null argument");
    if (!n) return 0;
    CUDA_TRY(cudaSetDevice(idx->device));
    DevBuf d_s, d_n, d_h;
    int rc;
    const size_t L = idx->dev.seed_len;
    do {
        if ((rc = d_s.ensure(n * L)) || (rc = d_n.ensure((size_t)n * 8)) || (rc = d_h.ensure((size_t)n * 2 * std::max(max_out, 1u) * 4))) break;
        cudaMemcpyAsync(d_s.p, seeds, n * L, cudaMemcpyHostToDevice, idx->stream);
        cudaMemsetAsync(d_h.p, 0, (size_t)n * 2 * std::max(max_out, 1u) * 4, idx->stream);
        lookup_kernel<<<(n + 127) / 128, 128, 0, idx->stream>>>(idx->dev, n, d_s.as<uint8_t>(), max_out, d_n.as<uint32_t>(), d_h.as<uint32_t>());
        cudaMemcpyAsync(n_hits, d_n.p, (size_t)n * 8, cudaMemcpyDeviceToHost, idx->stream);
        if (max_out) cudaMemcpyAsync(hits, d_h.p, (size_t)n * 2 * max_out * 4, cudaMemcpyDeviceToHost, idx->stream);
        cudaError_t e = cudaStreamSynchronize(idx->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "lookup_kernel: %s", cudaGetErrorString(e)); break; }
    } while (0);
    d_s.release(); d_n.release(); d_h.release();
    return rc;
}

extern "C" int snapb200_mapq_batch(int device, uint32_t n, const double *p_all, const double *p_best, const int32_t *score,
                                   const int32_t *popular_seeds_skipped, int32_t *mapq)
{
    return snapb200_mapq_batch_ex(device, n, p_all, p_best, score, popular_seeds_skipped, mapq, nullptr);
}

extern "C" int snapb200_mapq_batch_ex(int device, uint32_t n, const double *p_all, const double *p_best, const int32_t *score,
                                      const int32_t *popular_seeds_skipped, int32_t *mapq, uint8_t *host_reevaluated)
{
    snapb200_index *x;
    int rc = tables_for(device, &x);
    if (rc) return rc;
    if (!n) return 0;
    if (!p_all || !p_best || !score || !popular_seeds_skipped || !mapq) return set_error(SNAPB200_ERR_ARG, "null argument");
    DevBuf a, b, c, d, o, f;
    std::vector<uint8_t> flags(n);
    do {
        if ((rc = a.ensure((size_t)n * 8)) || (rc = b.ensure((size_t)n * 8)) || (rc = c.ensure((size_t)n * 4)) || (rc = d.ensure((size_t)n * 4)) ||
            (rc = o.ensure((size_t)n * 4)) || (rc = f.ensure(n))) break;
        cudaMemcpyAsync(a.p, p_all, (size_t)n * 8, cudaMemcpyHostToDevice, x->stream);
        cudaMemcpyAsync(b.p, p_best, (size_t)n * 8, cudaMemcpyHostToDevice, x->stream);
        cudaMemcpyAsync(c.p, score, (size_t)n * 4, cudaMemcpyHostToDevice, x->stream);
        cudaMemcpyAsync(d.p, popular_seeds_skipped, (size_t)n * 4, cudaMemcpyHostToDevice, x->stream);
        mapq_kernel<<<(n + 255) / 256, 256, 0, x->stream>>>(n, a.as<double>(), b.as<double>(), c.as<int32_t>(), d.as<int32_t>(), o.as<int32_t>(), f.as<uint8_t>());
        cudaMemcpyAsync(mapq, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost, x->stream);
        cudaMemcpyAsync(flags.data(), f.p, n, cudaMemcpyDeviceToHost, x->stream);
        cudaError_t e = cudaStreamSynchronize(x->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "mapq_kernel: %s", cudaGetErrorString(e)); break; }
        for (uint32_t i = 0; i < n; i++) if (flags[i]) mapq[i] = compute_mapq_host(p_all[i], p_best[i], score[i], popular_seeds_skipped[i]);
        if (host_reevaluated) memcpy(host_reevaluated, flags.data(), n);
    } while (0);
    DevBuf *all[] = {&a, &b, &c, &d, &o, &f};
    for (DevBuf *q : all) q->release();
    return rc;
}

// ---- stage-2 roofline diagnostics ------------------------------------------------------------------------------------------
extern "C" int snapb200_probe_bench(snapb200_index *idx, uint32_t n, const uint32_t *positions, uint32_t iters, float *ms_per_pass,
                                    uint64_t *slots_examined, uint64_t *count_words, uint64_t *hits_reported)
{
    if (!idx || !positions || !n || !iters || !ms_per_pass) return set_error(SNAPB200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    for (uint32_t i = 0; i < n; i++)
        if ((uint64_t)positions[i] + idx->dev.seed_len > idx->dev.n_bases) return set_error(SNAPB200_ERR_ARG, "position %u out of range", positions[i]);
    DevBuf d_pos, d_tot, d_packed;
    int rc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if ((rc = d_pos.ensure((size_t)n * 4)) || (rc = d_tot.ensure(32)) || (rc = d_packed.ensure((size_t)n * 16))) break;
        cudaMemcpyAsync(d_pos.p, positions, (size_t)n * 4, cudaMemcpyHostToDevice, idx->stream);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int T = 256;
        probe_pack_kernel<<<(n + T - 1) / T, T, 0, idx->stream>>>(idx->dev, n, d_pos.as<uint32_t>(), d_packed.as<ulonglong2>());
        probe_bench_kernel<<<(n + T - 1) / T, T, 0, idx->stream>>>(idx->dev, n, d_packed.as<ulonglong2>(), d_tot.as<unsigned long long>());  // warm-up
        cudaMemsetAsync(d_tot.p, 0, 32, idx->stream);
        cudaEventRecord(e0, idx->stream);
        for (uint32_t it = 0; it < iters; it++)
            probe_bench_kernel<<<(n + T - 1) / T, T, 0, idx->stream>>>(idx->dev, n, d_packed.as<ulonglong2>(), d_tot.as<unsigned long long>());
        cudaEventRecord(e1, idx->stream);
        cudaError_t e = cudaStreamSynchronize(idx->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "probe_bench_kernel: %s", cudaGetErrorString(e)); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_pass = ms / iters;
        unsigned long long tot[4];
        cudaMemcpy(tot, d_tot.p, 32, cudaMemcpyDeviceToHost);
        if (slots_examined) *slots_examined = tot[0] / iters;
        if (count_words) *count_words = tot[1] / iters;
        if (hits_reported) *hits_reported = tot[2] / iters;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    d_pos.release(); d_tot.release(); d_packed.release();
    return rc;
}

extern "C" int snapb200_gather_bench(int device, uint64_t bytes, uint32_t n, uint32_t iters, float *ms_per_pass)
{
    snapb200_index *x;
    int rc = tables_for(device, &x);
    if (rc) return rc;
    if (bytes < 4096 || !n || !iters || !ms_per_pass) return set_error(SNAPB200_ERR_ARG, "bad argument");
    DevBuf buf, sink;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if ((rc = buf.ensure(bytes)) || (rc = sink.ensure(8))) break;
        cudaMemsetAsync(buf.p, 0x5a, bytes, x->stream);
        cudaMemsetAsync(sink.p, 0, 8, x->stream);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int T = 256;
        const unsigned long long n_sectors = bytes / 32;
        gather_bench_kernel<<<(n + T - 1) / T, T, 0, x->stream>>>(buf.as<uint4>(), n_sectors, n, 0, sink.as<unsigned long long>());
        cudaEventRecord(e0, x->stream);
        for (uint32_t it = 0; it < iters; it++)
            gather_bench_kernel<<<(n + T - 1) / T, T, 0, x->stream>>>(buf.as<uint4>(), n_sectors, n, it + 1, sink.as<unsigned long long>());
        cudaEventRecord(e1, x->stream);
        cudaError_t e = cudaStreamSynchronize(x->stream);
        if (e != cudaSuccess) { rc = set_error(SNAPB200_ERR_CUDA, "gather_bench_kernel: %s", cudaGetErrorString(e)); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_pass = ms / iters;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    buf.release(); sink.release();
    return rc;
}

// ---- statistics -----------------------------------------------------------------------------------------------------
extern "C" int snapb200_stats_get(snapb200_index *idx, snapb200_stats *out)
{
    if (!idx || !out) return set_error(SNAPB200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long w[SNAPB200_STATS_WORDS];
    CUDA_TRY(cudaMemcpy(w, idx->stats, sizeof(w), cudaMemcpyDeviceToHost));
    int64_t *o = (int64_t *)out;
    for (int i = 0; i < SNAPB200_STATS_WORDS; i++) o[i] = (int64_t)w[i];
    out->useful_reads = out->total_reads - out->n_reads_ignored_ns;
    out->lv_calls = out->n_locations_scored;
    return 0;
}

// AlignerStats::add over the per-GPU copies of an index (SNAPLib/AlignerStats.cpp:75-102 sums the per-thread objects): one host
// process drives all GPUs of the box, so the reduction is a host sum; processes on several boxes reduce this flat int64 vector with
// one all-reduce (bench.py does, over NCCL).
extern "C" int snapb200_stats_sum(snapb200_index *const *indices, uint32_t n, snapb200_stats *out)
{
    if (!out || (n && !indices)) return set_error(SNAPB200_ERR_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    int64_t *o = (int64_t *)out;
    for (uint32_t i = 0; i < n; i++) {
        snapb200_stats one;
        int rc = snapb200_stats_get(indices[i], &one);
        if (rc) return rc;
        const int64_t *w = (const int64_t *)&one;
        for (int k = 0; k < SNAPB200_STATS_WORDS; k++) o[k] += w[k];
    }
    return 0;
}

extern "C" int snapb200_stats_reset(snapb200_index *idx)
{
    if (!idx) return set_error(SNAPB200_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(idx->device));
    CUDA_TRY(cudaMemset(idx->stats, 0, SNAPB200_STATS_WORDS * 8));
    return 0;
}

#include "io_api.inl"
#include "filter_api.inl"
