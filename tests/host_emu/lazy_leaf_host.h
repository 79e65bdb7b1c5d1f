// TEST INFRASTRUCTURE (included by csrc/lazy.cuh only when compiled as plain C++ for tests/host_emu): portable versions of the leaf
// primitives that csrc/fp_ptx.cuh implements in PTX, with the overflow / borrow checks that the PTX simulator of tools/gen_fp_ptx.py
// asserts, and a one-thread stand-in for the shared-memory slots.  Never part of the shipped library.
// (<cstdio> / <cstdlib> are included by lazy.cuh before it opens its namespace)

#define LZ_FN static
#define LZ_FN2 static
#define LZ_INL static inline
struct lz_u4 { uint32_t x, y, z, w; };
static lz_u4 lz_sm[2 * 64 * LZ_NT];
static inline uint32_t lz_tid() { return 0; }
static inline void lz_ld(uint32_t* r, uint32_t idx) {
    lz_u4 a = lz_sm[idx], b = lz_sm[idx + LZ_NT];
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
static inline void lz_st(uint32_t idx, const uint32_t* r) {
    lz_sm[idx] = lz_u4{r[0], r[1], r[2], r[3]}; lz_sm[idx + LZ_NT] = lz_u4{r[4], r[5], r[6], r[7]};
}
static inline void lz_die(const char* what) { fprintf(stderr, "lazy leaf check failed: %s\n", what); abort(); }

static inline void lz_mulw(uint32_t* w, const uint32_t* a, const uint32_t* b) {
    uint32_t t[16] = {0};
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)a[j] * b[i] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        t[i + 8] = (uint32_t)c;
    }
    for (int i = 0; i < 16; i++) w[i] = t[i];
}
static inline void lz_chain(uint32_t* r, const uint32_t* a, const uint32_t* b, int n, int sub, const char* what) {
    uint64_t c = 0; uint32_t t[16];
    for (int i = 0; i < n; i++) {
        if (sub) { uint64_t d = (uint64_t)a[i] - b[i] - c; t[i] = (uint32_t)d; c = (d >> 63) & 1; }
        else { c += (uint64_t)a[i] + b[i]; t[i] = (uint32_t)c; c >>= 32; }
    }
    if (c) lz_die(what);
    for (int i = 0; i < n; i++) r[i] = t[i];
}
static inline void lz_add8(uint32_t* r, const uint32_t* a, const uint32_t* b) { lz_chain(r, a, b, 8, 0, "add8 carry"); }
static inline void lz_sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) { lz_chain(r, a, b, 8, 1, "sub8 borrow"); }
static inline void lz_addw(uint32_t* r, const uint32_t* a, const uint32_t* b) { lz_chain(r, a, b, 16, 0, "addw carry"); }
static inline void lz_subw(uint32_t* r, const uint32_t* a, const uint32_t* b) { lz_chain(r, a, b, 16, 1, "subw borrow"); }
static inline void lz_addhi(uint32_t* r, const uint32_t* x, const uint32_t* c) {
    uint32_t t[16]; uint64_t cy = 0;
    for (int i = 0; i < 7; i++) t[i] = x[i];
    for (int i = 0; i < 9; i++) { cy += (uint64_t)x[7 + i] + c[i]; t[7 + i] = (uint32_t)cy; cy >>= 32; }
    if (cy) lz_die("addhi carry");
    for (int i = 0; i < 16; i++) r[i] = t[i];
}
static inline void lz_hi_chain(uint32_t* r, const uint32_t* x, const uint32_t* n, int sub, const char* what) {
    uint32_t t[16]; uint64_t c = 0;
    for (int i = 0; i < 8; i++) t[i] = x[i];
    for (int i = 0; i < 8; i++) {
        if (sub) { uint64_t d = (uint64_t)x[8 + i] - n[i] - c; t[8 + i] = (uint32_t)d; c = (d >> 63) & 1; }
        else { c += (uint64_t)x[8 + i] + n[i]; t[8 + i] = (uint32_t)c; c >>= 32; }
    }
    if (c) lz_die(what);
    for (int i = 0; i < 16; i++) r[i] = t[i];
}
static inline void lz_addw_hi(uint32_t* r, const uint32_t* x, const uint32_t* n) { lz_hi_chain(r, x, n, 0, "addw_hi carry"); }
static inline void lz_subw_hi(uint32_t* r, const uint32_t* x, const uint32_t* n) { lz_hi_chain(r, x, n, 1, "subw_hi borrow"); }
static inline void lz_shl3w(uint32_t* r, const uint32_t* x) {
    if (x[15] >> 29) lz_die("shl3w overflow");
    uint32_t t[16];
    t[0] = x[0] << 3;
    for (int i = 1; i < 16; i++) t[i] = (x[i] << 3) | (x[i - 1] >> 29);
    for (int i = 0; i < 16; i++) r[i] = t[i];
}
static inline void lz_csub_top(uint32_t* r, const uint32_t* x, const uint32_t* k, int n) {
    uint32_t t[8]; uint64_t bo = 0; int lo = n - 8;
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)x[lo + i] - k[i] - bo; t[i] = (uint32_t)d; bo = (d >> 63) & 1; }
    for (int i = 0; i < lo; i++) r[i] = x[i];
    for (int i = 0; i < 8; i++) r[lo + i] = bo ? x[lo + i] : t[i];
}
static inline void lz_csubw(uint32_t* r, const uint32_t* x, const uint32_t* k) { lz_csub_top(r, x, k, 16); }
static inline void lz_csub8(uint32_t* r, const uint32_t* x, const uint32_t* k) { lz_csub_top(r, x, k, 8); }
// r = (T + m p) / 2^256 without the final subtraction; the caller guarantees T < 4 p 2^256 (checked: no carry out of 256 bits)
static inline void lz_redc(uint32_t* r, const uint32_t* w) {
    uint32_t t[18];
    for (int i = 0; i < 16; i++) t[i] = w[i];
    t[16] = t[17] = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * 0xe4866389u; uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)m * C_P[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        for (int j = i + 8; c && j < 18; j++) { c += t[j]; t[j] = (uint32_t)c; c >>= 32; }
    }
    if (t[16] || t[17]) lz_die("redc result exceeds 256 bits");
    for (int i = 0; i < 8; i++) r[i] = t[8 + i];
}
static inline void fp_add_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fp x, y, z; for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; } fp_add(z, x, y); for (int i = 0; i < 8; i++) r[i] = z.v[i]; }
static inline void fp_sub_ptx(uint32_t* r, const uint32_t* a, const uint32_t* b) { fp x, y, z; for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; } fp_sub(z, x, y); for (int i = 0; i < 8; i++) r[i] = z.v[i]; }
