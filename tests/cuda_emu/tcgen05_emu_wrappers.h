// The inline-PTX wrappers of aecf_b200/csrc/gemm_tcgen05.cu on the host emulation -- TEST INFRASTRUCTURE ONLY.
// Included by that file INSIDE namespace aecf::tc (same names and signatures as the PTX versions); the model behind
// them is described in tcgen05_emu.h.
inline uint32_t smem_u32(const void* p) { return ::cuda_emu::tc::shared_address(p); }
inline void mbar_init(uint64_t* bar, uint32_t count) { ::cuda_emu::tc::mbar_init(bar, count); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { ::cuda_emu::tc::mbar_arrive(bar, -1, bytes); }
inline void mbar_arrive(uint64_t* bar) { ::cuda_emu::tc::mbar_arrive(bar, -1, 0); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (!::cuda_emu::tc::mbar_test(bar, parity)) ::cuda_emu::yield(); }
inline void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
inline void mbar_arrive_on_cta(uint64_t* bar, uint32_t target_cta) { ::cuda_emu::tc::mbar_arrive(bar, static_cast<int>(target_cta), 0); }
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void cluster_sync_all() { ::cuda_emu::cluster_barrier(); }
inline uint32_t cluster_ctarank() { return static_cast<uint32_t>(::cuda_emu::cta_rank()); }
inline bool elect_one() { return ::cuda_emu::thread().lane == 0; }

inline void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    ::cuda_emu::tc::tma_load(map, bar, -1, dst, c0, c1, 1u << ::cuda_emu::cta_rank());
}
inline void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    ::cuda_emu::tc::tma_load(map, bar, -1, dst, c0, c1, mask);     // lands, and completes bytes, in every CTA of the mask
}
inline void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    ::cuda_emu::tc::tma_load(map, bar, 0, dst, c0, c1, 1u << ::cuda_emu::cta_rank());   // bytes complete on the LEADER's barrier
}
inline void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) { ::cuda_emu::tc::tma_store(map, src, c0, c1); }
inline void tma_store_commit() { ::cuda_emu::tc::store_commit(); }
inline void tma_store_wait_read() { ::cuda_emu::tc::store_wait_read(0); }
inline void tma_store_wait_all() { ::cuda_emu::tc::store_wait_read(0); }
inline void prefetch_tmap(const CUtensorMap*) {}

inline void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    ::cuda_emu::tc::mma_f16(1, tmem_d, a_desc, b_desc, idesc, acc != 0);
}
inline void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    ::cuda_emu::tc::mma_f16(2, tmem_d, a_desc, b_desc, idesc, acc != 0);
}
inline void umma_commit(uint64_t* bar) { ::cuda_emu::tc::mbar_arrive(bar, -1, 0); }      // the MMAs above have already "retired"
inline void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    for (int c = 0; c < ::cuda_emu::cluster_size(); ++c)
        if (mask & (1u << c)) ::cuda_emu::tc::mbar_arrive(bar, c, 0);
}
inline void umma_commit_2sm(uint64_t* bar, uint16_t mask) { umma_commit_mc(bar, mask); }
inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) { ::cuda_emu::tc::tmem_load_32x32(taddr, r); }
inline void tmem_ld_wait() { ::cuda_emu::tc::tmem_load_wait(); }
inline void tmem_ld_wait_on(uint32_t (&)[32]) { ::cuda_emu::tc::tmem_load_wait(); }
