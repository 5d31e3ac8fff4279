// TEST INFRASTRUCTURE — host emulation of the PTX wrappers of word2bits_b200/csrc/w2b_ring.cuh (included from
// there, inside namespace w2b, only when W2B_EMULATE is defined).  Shared memory is the flat buffer w2b::smem;
// a "shared address" is an offset into it.  mbarriers, TMA bulk copies and bulk reduces keep their semantics:
// a copy lands (and counts its bytes on its mbarrier) at some later scheduling point, a bulk reduce reads its
// shared-memory source only when it is executed — at the latest when a wait_group.read of the issuing thread
// forces it — so a slot that is recycled too early corrupts the result here exactly as it would on the GPU.
extern unsigned char smem[];

void emu_mbar_init(unsigned off, int count);
void emu_mbar_expect_tx(unsigned off, unsigned bytes);
bool emu_mbar_try_wait(unsigned off, unsigned parity);
void emu_bulk_load(unsigned dst_off, const void *src, unsigned bytes, unsigned bar_off);
void emu_bulk_reduce_add(void *dst, unsigned src_off, unsigned bytes);
void emu_bulk_commit();
void emu_bulk_wait(int keep_groups);
// every checked shared-memory access lies inside the planned carve-up and, in the row-structured regions
// (u-ring, v-ring, staging rows, context_avg, error partials), inside ONE row: a column index that runs past
// the row would read or clobber the neighbouring slot on the GPU without any fault
void emu_check_smem(unsigned off, unsigned bytes, const char *what);
// cross-proxy rules of the bulk (TMA) engine, which reads and writes shared memory through the async proxy:
// a generic-proxy store must be followed by the storing thread's fence.proxy.async before a bulk reduce may
// read those bytes, and nothing may overwrite the source of a bulk reduce that has not been confirmed read
void emu_generic_store(unsigned off, unsigned bytes);
void emu_fence_async();

inline unsigned smem_u32(const void *p) { return (unsigned)((const unsigned char *)p - smem); }
inline void mbar_init(unsigned long long *bar, int count) { emu_mbar_init(smem_u32(bar), count); }
inline void mbar_expect_tx(unsigned bar, unsigned bytes) { emu_mbar_expect_tx(bar, bytes); }
inline void mbar_wait(unsigned bar, unsigned parity) {
  while (!emu_mbar_try_wait(bar, parity)) emu_yield();
}
inline void bulk_load(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar) {
  emu_bulk_load(dst_smem, src, bytes, bar);
}
inline void bulk_reduce_add(void *dst, unsigned src_smem, unsigned bytes) { emu_bulk_reduce_add(dst, src_smem, bytes); }
inline void bulk_commit() { emu_bulk_commit(); }
template <int N>
inline void bulk_wait_read() { emu_bulk_wait(N); }
inline void bulk_wait_all() { emu_bulk_wait(0); }
inline void fence_async_smem() { emu_fence_async(); }
inline void consumer_bar(int nthreads) { emu_block_barrier(1, nthreads); }
inline float4 lds128(unsigned addr) {
  float4 v;
  emu_check_smem(addr, 16, "ld.shared.v4");
  memcpy(&v, smem + addr, sizeof v);
  return v;
}
inline void sts128(unsigned addr, float4 v) {
  emu_check_smem(addr, 16, "st.shared.v4");
  emu_generic_store(addr, 16);
  memcpy(smem + addr, &v, sizeof v);
}
