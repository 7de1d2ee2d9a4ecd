// State and semantics behind tcgen05_emu.h -- TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"
#include "tcgen05_emu.h"

#include <cuda_bf16.h>

#include <vector>

namespace cuda_emu {
namespace tc {
namespace {

struct MBar { int expected, pending; long long tx; uint32_t phase; bool live; };
struct TensorMap {                       // what encode_tiled keeps inside the 128 opaque bytes of a CUtensorMap
    uint64_t magic;
    unsigned char* base;
    uint64_t inner, outer, row_stride_bytes;
    uint32_t element_bytes, box_inner, box_outer;
};
constexpr uint64_t MAGIC = 0x414543465f544d41ull;
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "the emulated tensor map must fit in CUtensorMap");

MBar g_bars[MAX_CLUSTER][SMEM_BYTES / 8];
float g_tmem[MAX_CLUSTER][128][512];

// bulk async-groups are per thread: stores issued, groups committed, not yet read out of shared memory
struct PendingStore { TensorMap map; int cta; size_t src_off; int c0, c1; };
struct StoreQueue { std::vector<PendingStore> open; std::vector<std::vector<PendingStore>> committed; };
StoreQueue g_stores[MAX_CLUSTER][1024];

[[noreturn]] void fail(const char* what) {
    std::fprintf(stderr, "cuda_emu/tcgen05: %s\n", what);
    std::abort();
}

// (CTA of the running cluster, byte offset) of a pointer into shared memory
void locate(const void* p, int* cta, size_t* offset) {
    const unsigned char* q = static_cast<const unsigned char*>(p);
    for (int c = 0; c < MAX_CLUSTER; ++c) {
        const unsigned char* w = smem_window(c);
        if (q >= w && q < w + SMEM_BYTES) { *cta = c; *offset = static_cast<size_t>(q - w); return; }
    }
    fail("pointer is not in shared memory");
}

MBar& bar_at(const void* bar, int target_cta) {
    int cta; size_t off;
    locate(bar, &cta, &off);
    if (off % 8 != 0) fail("misaligned mbarrier");
    return g_bars[target_cta >= 0 ? target_cta : cta][off / 8];
}

void settle(MBar& b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.expected; made_progress(); }
    if (b.pending < 0) fail("more arrivals than the mbarrier was initialised for");
}

// 128-byte swizzle: bits [4, 7) of a shared-memory byte address are XORed with bits [7, 10)
inline size_t swizzle128(size_t address) { return address ^ (((address >> 7) & 7) << 4); }

const TensorMap& view(const CUtensorMap* map) {
    const TensorMap& t = *reinterpret_cast<const TensorMap*>(map);
    if (t.magic != MAGIC) fail("tensor map was not made by the emulated cuTensorMapEncodeTiled");
    return t;
}

float bf16_at(int cta, size_t address) {
    uint16_t bits;
    std::memcpy(&bits, smem_window(cta) + swizzle128(address), 2);
    uint32_t wide = static_cast<uint32_t>(bits) << 16;
    float f;
    std::memcpy(&f, &wide, 4);
    return f;
}

struct Operand { size_t start, lbo, sbo; };
Operand decode(uint64_t desc) {
    if (((desc >> 61) & 7) != 2) fail("shared-memory descriptor is not SWIZZLE_128B");
    return {static_cast<size_t>(desc & 0x3FFF) << 4, static_cast<size_t>((desc >> 16) & 0x3FFF) << 4,
            static_cast<size_t>((desc >> 32) & 0x3FFF) << 4};
}
// element (mn, k) of a 16-deep operand slice
//   K-major : rows of 128 B along K, 8-row groups SBO apart
//   MN-major: 64-wide MN atoms LBO apart, each [K rows x 128 B] with 8-K-row groups SBO apart
float operand_at(int cta, const Operand& o, bool mn_major, int mn, int k) {
    const size_t address = mn_major ? o.start + static_cast<size_t>(mn / 64) * o.lbo + static_cast<size_t>(k / 8) * o.sbo + (k % 8) * 128 + (mn % 64) * 2
                                    : o.start + static_cast<size_t>(mn / 8) * o.sbo + (mn % 8) * 128 + k * 2;
    return bf16_at(cta, address);
}

}  // namespace

uint32_t shared_address(const void* p) {
    int cta; size_t off;
    locate(p, &cta, &off);
    return static_cast<uint32_t>(off) | (static_cast<uint32_t>(cta) << 24);
}

void mbar_init(const void* bar, uint32_t count) {
    MBar& b = bar_at(bar, -1);
    b = MBar{static_cast<int>(count), static_cast<int>(count), 0, 0u, true};
}

void mbar_arrive(const void* bar, int target_cta, uint32_t expect_tx_bytes) {
    MBar& b = bar_at(bar, target_cta);
    if (!b.live) fail("arrive on an mbarrier that was never initialised");
    b.tx += expect_tx_bytes;
    --b.pending;
    settle(b);
}

bool mbar_test(const void* bar, uint32_t parity) {
    const MBar& b = bar_at(bar, -1);
    if (!b.live) fail("wait on an mbarrier that was never initialised");
    return b.phase != (parity & 1u);                     // the phase of that parity has completed
}

void tma_load(const CUtensorMap* map, const void* bar, int bar_cta, void* dst, int c0, int c1, unsigned cta_mask) {
    const TensorMap& t = view(map);
    int own; size_t dst_off, bar_off;
    locate(dst, &own, &dst_off);
    locate(bar, &own, &bar_off);
    if (t.box_inner * t.element_bytes != 128 || dst_off % 1024 != 0) fail("TMA box is not one 1024-aligned 128-byte swizzle atom wide");
    const size_t bytes = static_cast<size_t>(t.box_outer) * 128;
    for (int c = 0; c < cluster_size(); ++c) {
        if (!(cta_mask & (1u << c))) continue;
        unsigned char* window = smem_window(c);
        for (uint32_t r = 0; r < t.box_outer; ++r)
            for (uint32_t i = 0; i < t.box_inner; ++i) {
                const long long row = static_cast<long long>(c1) + r, col = static_cast<long long>(c0) + i;
                unsigned char* to = window + swizzle128(dst_off + static_cast<size_t>(r) * 128 + static_cast<size_t>(i) * t.element_bytes);
                if (row >= 0 && col >= 0 && row < static_cast<long long>(t.outer) && col < static_cast<long long>(t.inner))
                    std::memcpy(to, t.base + row * t.row_stride_bytes + col * t.element_bytes, t.element_bytes);
                else
                    std::memset(to, 0, t.element_bytes);          // out of bounds: zero fill
            }
        MBar& b = g_bars[bar_cta >= 0 ? bar_cta : c][bar_off / 8];
        if (!b.live) fail("TMA load signals an mbarrier that was never initialised");
        b.tx -= static_cast<long long>(bytes);                    // the whole box counts, in bounds or not
        settle(b);
    }
}

namespace {
void perform_store(const PendingStore& ps) {
    const TensorMap& t = ps.map;
    const int cta = ps.cta;
    const size_t src_off = ps.src_off;
    const int c0 = ps.c0, c1 = ps.c1;
    const unsigned char* window = smem_window(cta);
    for (uint32_t r = 0; r < t.box_outer; ++r)
        for (uint32_t i = 0; i < t.box_inner; ++i) {
            const long long row = static_cast<long long>(c1) + r, col = static_cast<long long>(c0) + i;
            if (row < 0 || col < 0 || row >= static_cast<long long>(t.outer) || col >= static_cast<long long>(t.inner)) continue;   // clipped
            std::memcpy(t.base + row * t.row_stride_bytes + col * t.element_bytes,
                        window + swizzle128(src_off + static_cast<size_t>(r) * 128 + static_cast<size_t>(i) * t.element_bytes), t.element_bytes);
        }
}
StoreQueue& my_stores() { return g_stores[cta_rank()][thread().linear]; }
}  // namespace

void tma_store(const CUtensorMap* map, const void* src, int c0, int c1) {
    const TensorMap& t = view(map);
    int cta; size_t src_off;
    locate(src, &cta, &src_off);
    if (t.box_inner * t.element_bytes != 128 || src_off % 1024 != 0) fail("TMA store box is not one 1024-aligned 128-byte swizzle atom wide");
    my_stores().open.push_back(PendingStore{t, cta, src_off, c0, c1});      // shared memory is read when the group retires
}

void store_commit() {
    StoreQueue& q = my_stores();
    q.committed.push_back(std::move(q.open));
    q.open.clear();
}

void store_wait_read(int pending_allowed) {
    StoreQueue& q = my_stores();
    while (static_cast<int>(q.committed.size()) > pending_allowed) {
        for (const PendingStore& ps : q.committed.front()) perform_store(ps);
        q.committed.erase(q.committed.begin());
    }
}

void tmem_alloc(uint32_t* slot, uint32_t columns) {
    if (columns > 512 || (columns & (columns - 1)) != 0 || columns < 32) fail("tcgen05.alloc: column count must be a power of two in [32, 512]");
    *slot = 0;                                           // lane 0, column 0: one allocation per CTA in these kernels
}

void mma_f16(int cta_group, uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const int n = static_cast<int>((idesc >> 17) & 0x3F) << 3, m = static_cast<int>((idesc >> 24) & 0x1F) << 4;
    const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
    if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 1 || ((idesc >> 10) & 7) != 1) fail("instruction descriptor is not bf16 x bf16 -> f32");
    if (m != 128 * cta_group || n < 16 || n > 256 || n % 16 != 0) fail("unsupported MMA shape");
    const Operand a = decode(a_desc), b = decode(b_desc);
    const int col0 = static_cast<int>(tmem_d & 0xFFFF), lane0 = static_cast<int>(tmem_d >> 16);
    if (lane0 != 0 || col0 + n > 512) fail("accumulator outside tensor memory");
    const int leader = cuda_emu::cta_rank();
    if (cta_group == 2 && leader != 0) fail("cta_group::2 MMA issued by the non-leader CTA");
    const int n_per_cta = n / cta_group;                 // cta_group::2: each CTA stages its own half of the B columns
    for (int g = 0; g < cta_group; ++g) {
        const int cta = cta_group == 2 ? g : leader;     // rows 128 g .. of the tile come from (and go to) CTA g
        for (int r = 0; r < 128; ++r) {
            float a_row[16];
            for (int k = 0; k < 16; ++k) a_row[k] = operand_at(cta, a, a_mn, r, k);
            for (int j = 0; j < n; ++j) {
                const int b_cta = cta_group == 2 ? j / n_per_cta : leader;
                float acc = 0.f;
                for (int k = 0; k < 16; ++k) acc += a_row[k] * operand_at(b_cta, b, b_mn, j % n_per_cta, k);
                float& d = g_tmem[cta][r][col0 + j];
                d = accumulate ? d + acc : acc;
            }
        }
    }
}

// tcgen05.ld is asynchronous: the destination registers hold the data only after tcgen05.wait::ld.  Until then they are
// poisoned here (a quiet NaN pattern), and the values are delivered by the wait.
namespace {
struct PendingLoad { uint32_t* dst; uint32_t values[32]; };
std::vector<PendingLoad> g_loads[MAX_CLUSTER][1024];
}  // namespace

void tmem_load_32x32(uint32_t taddr, uint32_t* out32) {
    const int lane_base = static_cast<int>(taddr >> 16), col = static_cast<int>(taddr & 0xFFFF);
    const ThreadState& t = thread();
    if (lane_base != (t.warp % 4) * 32) fail("tcgen05.ld: a warp may only read the TMEM lane quadrant warp_id % 4");
    if (col + 32 > 512) fail("tcgen05.ld beyond column 511");
    PendingLoad p;
    p.dst = out32;
    std::memcpy(p.values, &g_tmem[cta_rank()][lane_base + t.lane][col], 32 * sizeof(float));
    for (int i = 0; i < 32; ++i) out32[i] = 0x7FC0DEADu;
    g_loads[cta_rank()][t.linear].push_back(p);
}

void tmem_load_wait() {
    auto& q = g_loads[cta_rank()][thread().linear];
    for (const PendingLoad& p : q) std::memcpy(p.dst, p.values, sizeof(p.values));
    q.clear();
}

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave,
                      CUtensorMapSwizzle swizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (rank != 2 || swizzle != CU_TENSOR_MAP_SWIZZLE_128B) return CUDA_ERROR_INVALID_VALUE;
    const uint32_t es = dt == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 0;
    if (es == 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) || strides[0] % 16 != 0 || box[0] > 256 || box[1] > 256 || box[0] * es > 128)
        return CUDA_ERROR_INVALID_VALUE;                 // the constraints cuTensorMapEncodeTiled itself enforces
    std::memset(map, 0, sizeof(*map));
    TensorMap t{MAGIC, static_cast<unsigned char*>(ptr), dims[0], dims[1], strides[0], es, box[0], box[1]};
    std::memcpy(map, &t, sizeof(t));
    return CUDA_SUCCESS;
}

}  // namespace tc

void reset_block_resources(int cluster) {
    for (int c = 0; c < cluster; ++c) {
        for (auto& b : tc::g_bars[c]) b.live = false;
        for (auto& q : tc::g_stores[c])
            if (!q.open.empty() || !q.committed.empty()) tc::fail("the previous kernel exited with TMA stores it never waited for");
        for (auto& q : tc::g_loads[c]) q.clear();
    }
}

}  // namespace cuda_emu
