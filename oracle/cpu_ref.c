/* cpu_ref.c — C restatement of the reference's CPU loops.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * The reference (VGLoic/kzg-poly-commit-exploration) does all arithmetic through blst 0.3.15
 * (Cargo.toml:10), a third-party C/assembly library that is NOT under /root/reference and cannot be
 * built here, so this file restates the reference's *algorithms* over a portable 64-bit-limb
 * Montgomery implementation:
 *   commit    : acc = inf; for i: acc += [c_i] * setup[i]      src/polynomial.rs:207-212
 *               (one 256-bit windowed scalar multiplication + one add_or_double per coefficient,
 *                src/curves.rs:79-96; blst_p1_mult is a 5-bit signed-window ladder)
 *   evaluate  : sum c_i * pow(x, i) with the bespoke pow      src/polynomial.rs:113-118, src/scalar.rs:122-187
 *   open      : (p - y) / (x - z) synthetic division + commit   src/polynomial.rs:260-269, 150-195
 *   setup     : running s^i + one generator scalar-mul per item src/trusted_setup.rs:50-62
 * It is validated bit-exactly against oracle/kzg_oracle.py (tests/test_cpu_ref.py) and is timed by
 * bench.py as the `cpu_baseline` ("kind": "port").  Nothing in the product may link or call it.
 * "Parity pinned by spec + pairing": see the header of kzg_oracle.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ---- generic Montgomery arithmetic, N 64-bit limbs -------------------------------------------- */
#define NQ 6
#define NR 4
static const u64 Q[NQ] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                          0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 Q_INV = 0x89f3fffcfffcfffdULL; /* -q^-1 mod 2^64 */
static const u64 R_[NR] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                           0x73eda753299d7d48ULL};
static const u64 R_INV = 0xfffffffeffffffffULL; /* -r^-1 mod 2^64 */

typedef struct { u64 v[NQ]; } fq;
typedef struct { u64 v[NR]; } fr;

static void mont_mul(u64* o, const u64* a, const u64* b, const u64* p, u64 inv, int n) {
  u64 t[8] = {0};
  for (int i = 0; i < n; i++) {
    u128 c = 0;
    for (int j = 0; j < n; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
    c += t[n]; t[n] = (u64)c; t[n + 1] = (u64)(c >> 64);
    u64 m = t[0] * inv;
    c = ((u128)m * p[0] + t[0]) >> 64;
    for (int j = 1; j < n; j++) { c += (u128)m * p[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
    c += t[n]; t[n - 1] = (u64)c; t[n] = t[n + 1] + (u64)(c >> 64);
  }
  u64 u[6]; u64 br = 0;
  for (int j = 0; j < n; j++) { u128 d = (u128)t[j] - p[j] - br; u[j] = (u64)d; br = (u64)(d >> 64) & 1; }
  int ge = t[n] != 0 || br == 0;
  for (int j = 0; j < n; j++) o[j] = ge ? u[j] : t[j];
}
static void add_mod(u64* o, const u64* a, const u64* b, const u64* p, int n) {
  u64 t[6], u[6]; u128 c = 0; u64 br = 0;
  for (int j = 0; j < n; j++) { c += (u128)a[j] + b[j]; t[j] = (u64)c; c >>= 64; }
  for (int j = 0; j < n; j++) { u128 d = (u128)t[j] - p[j] - br; u[j] = (u64)d; br = (u64)(d >> 64) & 1; }
  int ge = c != 0 || br == 0;
  for (int j = 0; j < n; j++) o[j] = ge ? u[j] : t[j];
}
static void sub_mod(u64* o, const u64* a, const u64* b, const u64* p, int n) {
  u64 t[6]; u64 br = 0; u128 c = 0;
  for (int j = 0; j < n; j++) { u128 d = (u128)a[j] - b[j] - br; t[j] = (u64)d; br = (u64)(d >> 64) & 1; }
  for (int j = 0; j < n; j++) { c += (u128)t[j] + (br ? p[j] : 0); o[j] = (u64)c; c >>= 64; }
}

#define FQ_MUL(o, a, b) mont_mul((o).v, (a).v, (b).v, Q, Q_INV, NQ)
#define FQ_ADD(o, a, b) add_mod((o).v, (a).v, (b).v, Q, NQ)
#define FQ_SUB(o, a, b) sub_mod((o).v, (a).v, (b).v, Q, NQ)
#define FR_MUL(o, a, b) mont_mul((o).v, (a).v, (b).v, R_, R_INV, NR)
#define FR_ADD(o, a, b) add_mod((o).v, (a).v, (b).v, R_, NR)
#define FR_SUB(o, a, b) sub_mod((o).v, (a).v, (b).v, R_, NR)

static fq FQ_R2, FQ_ONE, FQ_ZERO; /* R^2 mod q, R mod q */
static fr FR_R2, FR_ONE, FR_ZERO;
static int g_init = 0;

static int is_zero(const u64* a, int n) { u64 x = 0; for (int i = 0; i < n; i++) x |= a[i]; return x == 0; }
static int eq(const u64* a, const u64* b, int n) { u64 x = 0; for (int i = 0; i < n; i++) x |= a[i] ^ b[i]; return x == 0; }

/* 2^k mod p by repeated doubling (used once, to derive R and R^2 without hard-coding them) */
static void pow2_mod(u64* o, int k, const u64* p, int n) {
  u64 t[6] = {1, 0, 0, 0, 0, 0};
  for (int i = 0; i < k; i++) add_mod(t, t, t, p, n);
  memcpy(o, t, 8 * n);
}
static void init_consts(void) {
  if (g_init) return;
  pow2_mod(FQ_ONE.v, 384, Q, NQ); pow2_mod(FQ_R2.v, 768, Q, NQ);
  pow2_mod(FR_ONE.v, 256, R_, NR); pow2_mod(FR_R2.v, 512, R_, NR);
  memset(&FQ_ZERO, 0, sizeof FQ_ZERO); memset(&FR_ZERO, 0, sizeof FR_ZERO);
  g_init = 1;
}

/* ---- Fr helpers: Scalar (src/scalar.rs) ---------------------------------------------------------- */
static void fr_from_le(fr* o, const uint8_t* b) { /* Scalar::from_le_bytes: reduce mod r (src/scalar.rs:54-61) */
  fr t; for (int i = 0; i < NR; i++) { u64 w = 0; for (int k = 7; k >= 0; k--) w = (w << 8) | b[8 * i + k]; t.v[i] = w; }
  FR_MUL(*o, FR_R2, t);
}
static void fr_to_le(uint8_t* b, const fr* a) { /* Scalar::to_le_bytes (src/scalar.rs:83-93) */
  fr one = {{1, 0, 0, 0}}, t; FR_MUL(t, *a, one);
  for (int i = 0; i < NR; i++) for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(t.v[i] >> (8 * k));
}
/* Scalar::pow — the reference's bespoke square-and-multiply, step by step (src/scalar.rs:122-187) */
static void fr_pow_ref(fr* o, const fr* x, u64 n) {
  if (n == 0) { *o = FR_ONE; return; }
  if (n == 1) { *o = *x; return; }
  u64 target = n / 2;
  fr powered[64]; int np = 0;
  u64 cur = 1;
  while (cur * 2 <= target) {
    const fr* last = np ? &powered[np - 1] : x;
    FR_MUL(powered[np], *last, *last); np++;
    cur *= 2;
  }
  fr acc; u64 tracker;
  if (target % 2 == 0) { acc = FR_ONE; tracker = 0; } else { acc = *x; tracker = 1; }
  u64 avail = 1ULL << np;
  while (tracker != target) {
    if (tracker + avail <= target) { tracker += avail; FR_MUL(acc, acc, powered[np - 1]); }
    np--; avail /= 2;
  }
  fr sq; FR_MUL(sq, acc, acc);
  if (target * 2 == n) *o = sq; else FR_MUL(*o, sq, *x);
}

/* ---- G1 Jacobian (what blst_p1 is): infinity <=> Z == 0 ------------------------------------------- */
typedef struct { fq x, y, z; } p1;

static void p1_set_inf(p1* p) { memset(p, 0, sizeof *p); }
static int p1_is_inf(const p1* p) { return is_zero(p->z.v, NQ); }

static void p1_double(p1* o, const p1* a) {
  if (p1_is_inf(a)) { *o = *a; return; }
  fq A, B, C, D, E, F, t;
  FQ_MUL(A, a->x, a->x); FQ_MUL(B, a->y, a->y); FQ_MUL(C, B, B);
  FQ_ADD(t, a->x, B); FQ_MUL(t, t, t); FQ_SUB(t, t, A); FQ_SUB(t, t, C); FQ_ADD(D, t, t);
  FQ_ADD(E, A, A); FQ_ADD(E, E, A);
  FQ_MUL(F, E, E);
  fq x3, y3, z3;
  FQ_SUB(x3, F, D); FQ_SUB(x3, x3, D);
  FQ_SUB(t, D, x3); FQ_MUL(y3, E, t);
  FQ_ADD(C, C, C); FQ_ADD(C, C, C); FQ_ADD(C, C, C); FQ_SUB(y3, y3, C);
  FQ_MUL(z3, a->y, a->z); FQ_ADD(z3, z3, z3);
  o->x = x3; o->y = y3; o->z = z3;
}
/* blst_p1_add_or_double: complete addition (src/curves.rs:82) */
static void p1_add(p1* o, const p1* a, const p1* b) {
  if (p1_is_inf(a)) { *o = *b; return; }
  if (p1_is_inf(b)) { *o = *a; return; }
  fq z1z1, z2z2, u1, u2, s1, s2, h, r, hh, hhh, v, t;
  FQ_MUL(z1z1, a->z, a->z); FQ_MUL(z2z2, b->z, b->z);
  FQ_MUL(u1, a->x, z2z2); FQ_MUL(u2, b->x, z1z1);
  FQ_MUL(s1, a->y, b->z); FQ_MUL(s1, s1, z2z2);
  FQ_MUL(s2, b->y, a->z); FQ_MUL(s2, s2, z1z1);
  if (eq(u1.v, u2.v, NQ)) {
    if (eq(s1.v, s2.v, NQ)) { p1_double(o, a); return; }
    p1_set_inf(o); return;
  }
  FQ_SUB(h, u2, u1); FQ_SUB(r, s2, s1);
  FQ_MUL(hh, h, h); FQ_MUL(hhh, h, hh); FQ_MUL(v, u1, hh);
  fq x3, y3, z3;
  FQ_MUL(x3, r, r); FQ_SUB(x3, x3, hhh); FQ_SUB(x3, x3, v); FQ_SUB(x3, x3, v);
  FQ_SUB(t, v, x3); FQ_MUL(y3, r, t); FQ_MUL(t, s1, hhh); FQ_SUB(y3, y3, t);
  FQ_MUL(z3, a->z, b->z); FQ_MUL(z3, z3, h);
  o->x = x3; o->y = y3; o->z = z3;
}
static void p1_neg(p1* p) { FQ_SUB(p->y, FQ_ZERO, p->y); }

/* blst_p1_mult(out, p, scalar_le, 256): 5-bit signed-window ladder over the canonical scalar bytes
 * (src/curves.rs:90-96).  51 windows + top, 16-entry table. */
static void p1_mult(p1* o, const p1* p, const uint8_t k_le[32]) {
  p1 tab[17];
  p1_set_inf(&tab[0]); tab[1] = *p;
  for (int i = 2; i <= 16; i++) { if (i & 1) p1_add(&tab[i], &tab[i - 1], p); else p1_double(&tab[i], &tab[i / 2]); }
  /* signed digits, 5 bits each, from the least significant end */
  int dig[53]; int carry = 0;
  for (int w = 0; w < 52; w++) {
    int bit = 5 * w, v = 0;
    for (int b = 0; b < 5; b++) { int pos = bit + b; if (pos < 256) v |= ((k_le[pos >> 3] >> (pos & 7)) & 1) << b; }
    v += carry;
    if (v > 16) { v -= 32; carry = 1; } else carry = 0;
    dig[w] = v;
  }
  dig[52] = carry;
  p1 acc; p1_set_inf(&acc);
  for (int w = 52; w >= 0; w--) {
    if (w != 52) for (int b = 0; b < 5; b++) p1_double(&acc, &acc);
    int d = dig[w];
    if (d > 0) p1_add(&acc, &acc, &tab[d]);
    else if (d < 0) { p1 t = tab[-d]; p1_neg(&t); p1_add(&acc, &acc, &t); }
  }
  *o = acc;
}

static void fq_inv(fq* o, const fq* a) { /* a^(q-2) */
  u64 e[NQ]; memcpy(e, Q, sizeof e); e[0] -= 2;
  fq acc = FQ_ONE;
  for (int i = NQ * 64 - 1; i >= 0; i--) {
    FQ_MUL(acc, acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) FQ_MUL(acc, acc, *a);
  }
  *o = acc;
}
/* blst_p1_compress (src/curves.rs:99-110) */
static void p1_compress(uint8_t out[48], const p1* p) {
  if (p1_is_inf(p)) { memset(out, 0, 48); out[0] = 0xC0; return; }
  fq zi, zi2, x, y, one = {{1, 0, 0, 0, 0, 0}};
  fq_inv(&zi, &p->z); FQ_MUL(zi2, zi, zi);
  FQ_MUL(x, p->x, zi2); FQ_MUL(y, p->y, zi2); FQ_MUL(y, y, zi);
  FQ_MUL(x, x, one); FQ_MUL(y, y, one);
  for (int i = 0; i < NQ; i++) for (int k = 0; k < 8; k++) out[47 - (8 * i + k)] = (uint8_t)(x.v[i] >> (8 * k));
  out[0] |= 0x80;
  /* y > (q-1)/2  <=>  2y > q-1  <=>  2y >= q+1 > q (y canonical) */
  u64 half[NQ]; u64 c = 0;
  for (int i = NQ - 1; i >= 0; i--) { u64 w = Q[i]; half[i] = (w >> 1) | (c << 63); c = w & 1; } /* (q-1)/2 since q odd */
  u64 br = 0;
  for (int i = 0; i < NQ; i++) { u128 d = (u128)half[i] - y.v[i] - br; br = (u64)(d >> 64) & 1; }
  if (br) out[0] |= 0x20;
}

static void generator(p1* g) {
  static const u64 gx[NQ] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                             0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
  static const u64 gy[NQ] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                             0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
  fq x, y; memcpy(x.v, gx, sizeof gx); memcpy(y.v, gy, sizeof gy);
  FQ_MUL(g->x, FQ_R2, x); FQ_MUL(g->y, FQ_R2, y); g->z = FQ_ONE;
}

/* ---- public entry points ---------------------------------------------------------------------------- */
typedef struct { u64 n; p1* g1; } ref_setup;

/* SetupArtifactsGenerator::new(secret).take(n) — G1 half (src/trusted_setup.rs:20-28, 40-62).
 * threads == 1 is the reference's behaviour; threads > 1 splits the index range (each range
 * seeds its running power with one square-and-multiply) — a parallelised restatement. */
typedef struct { ref_setup* s; fr secret; u64 lo, hi; } setup_job;
static void* setup_range(void* arg) {
  setup_job* j = (setup_job*)arg;
  p1 g; generator(&g);
  fr cur = FR_ONE, base = j->secret;
  for (u64 e = j->lo ? j->lo - 1 : 0; e; e >>= 1) { if (e & 1) FR_MUL(cur, cur, base); FR_MUL(base, base, base); }
  /* cur = s^(lo-1) (or 1 when lo == 0, where item 0 is handled separately) */
  for (u64 i = j->lo; i < j->hi; i++) {
    if (i == 0) { j->s->g1[0] = g; continue; }
    FR_MUL(cur, cur, j->secret);                   /* current_s_powered * secret, :50 */
    uint8_t kb[32]; fr_to_le(kb, &cur);
    p1_mult(&j->s->g1[i], &g, kb);               /* blst_p1_mult(generator, s^i, 256), :54-62 */
  }
  return NULL;
}
ref_setup* ref_setup_generate_mt(const uint8_t secret_be[32], u64 n, int threads) {
  init_consts();
  ref_setup* s = (ref_setup*)malloc(sizeof *s);
  s->n = n; s->g1 = (p1*)malloc(sizeof(p1) * (n ? n : 1));
  uint8_t le[32]; for (int i = 0; i < 32; i++) le[i] = secret_be[31 - i];
  fr secret; fr_from_le(&secret, le);
  if (threads < 1) threads = 1;
  if ((u64)threads > n) threads = n ? (int)n : 1;
  setup_job* jobs = (setup_job*)calloc(threads, sizeof *jobs);
  pthread_t* th = (pthread_t*)calloc(threads, sizeof *th);
  for (int t = 0; t < threads; t++) {
    jobs[t].s = s; jobs[t].secret = secret; jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
    if (threads == 1) setup_range(&jobs[t]); else pthread_create(&th[t], NULL, setup_range, &jobs[t]);
  }
  for (int t = 0; t < threads && threads > 1; t++) pthread_join(th[t], NULL);
  free(jobs); free(th);
  return s;
}
ref_setup* ref_setup_generate(const uint8_t secret_be[32], u64 n) { return ref_setup_generate_mt(secret_be, n, 1); }
void ref_setup_free(ref_setup* s) { if (s) { free(s->g1); free(s); } }
void ref_setup_export(const ref_setup* s, u64 i, uint8_t out[48]) { p1_compress(out, &s->g1[i]); }

typedef struct { const ref_setup* s; const fr* co; u64 lo, hi; p1 acc; } commit_job;
static void* commit_range(void* arg) {
  commit_job* j = (commit_job*)arg;
  p1 acc; p1_set_inf(&acc);
  for (u64 i = j->lo; i < j->hi; i++) { /* the hot loop, src/polynomial.rs:208-212 */
    uint8_t kb[32]; fr_to_le(kb, &j->co[i]);   /* G1Point::mult -> a.to_le_bytes() */
    p1 c; p1_mult(&c, &j->s->g1[i], kb);
    p1_add(&acc, &acc, &c);
  }
  j->acc = acc;
  return NULL;
}
static void commit_fr(const ref_setup* s, const fr* co, u64 n, int threads, p1* out) {
  if (threads < 1) threads = 1;
  if ((u64)threads > n) threads = n ? (int)n : 1;
  commit_job* jobs = (commit_job*)calloc(threads, sizeof *jobs);
  pthread_t* th = (pthread_t*)calloc(threads, sizeof *th);
  for (int t = 0; t < threads; t++) {
    jobs[t].s = s; jobs[t].co = co; jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
    if (threads == 1) commit_range(&jobs[t]); else pthread_create(&th[t], NULL, commit_range, &jobs[t]);
  }
  p1 acc; p1_set_inf(&acc);
  for (int t = 0; t < threads; t++) { if (threads > 1) pthread_join(th[t], NULL); p1_add(&acc, &acc, &jobs[t].acc); }
  free(jobs); free(th);
  *out = acc;
}
static u64 load_trimmed(fr** out, const uint8_t* coeffs, u64 n) { /* Polynomial::try_from (src/polynomial.rs:55-75) */
  fr* co = (fr*)malloc(sizeof(fr) * (n ? n : 1));
  u64 last = 0; int any = 0;
  for (u64 i = 0; i < n; i++) { fr_from_le(&co[i], coeffs + 32 * i); if (!is_zero(co[i].v, NR)) { last = i; any = 1; } }
  *out = co;
  return any ? last + 1 : 0;
}
/* Polynomial::commit; returns 0 ok, 1 degree too high.  threads == 1 is the reference's behaviour */
int ref_commit(const ref_setup* s, const uint8_t* coeffs_le32, u64 n, uint8_t out[48], int threads) {
  init_consts();
  fr* co; u64 nt = load_trimmed(&co, coeffs_le32, n);
  if ((nt ? nt : 1) > s->n) { free(co); return 1; }
  p1 acc; commit_fr(s, co, nt, threads, &acc);
  p1_compress(out, &acc);
  free(co);
  return 0;
}
/* Polynomial::evaluate with the reference's per-term pow (src/polynomial.rs:112-123) */
void ref_evaluate(const uint8_t* coeffs_le32, u64 n, const uint8_t z_le[32], uint8_t y_le[32]) {
  init_consts();
  fr* co; u64 nt = load_trimmed(&co, coeffs_le32, n);
  fr x, acc = FR_ZERO; fr_from_le(&x, z_le);
  for (u64 d = 0; d < nt; d++) { fr xp, c; fr_pow_ref(&xp, &x, d); FR_MUL(c, co[d], xp); FR_ADD(acc, acc, c); }
  fr_to_le(y_le, &acc);
  free(co);
}
/* Evaluation::generate_proof (src/polynomial.rs:260-269): 0 ok, 1 degree, 2 constant, 3 remainder */
int ref_open(const ref_setup* s, const uint8_t* coeffs_le32, u64 n, const uint8_t z_le[32], const uint8_t y_le[32],
             uint8_t out[48], int threads) {
  init_consts();
  fr* co; u64 nt = load_trimmed(&co, coeffs_le32, n);
  fr z, y; fr_from_le(&z, z_le); fr_from_le(&y, y_le);
  p1 acc; p1_set_inf(&acc);
  int rc = 0;
  if (nt <= 1) {
    fr c0 = nt ? co[0] : FR_ZERO, d; FR_SUB(d, c0, y);
    if (!is_zero(d.v, NR)) rc = 2;
  } else {
    FR_SUB(co[0], co[0], y);                     /* sub(from_constant(result)) */
    fr* q = (fr*)malloc(sizeof(fr) * (nt - 1));
    q[nt - 2] = co[nt - 1];
    for (u64 i = nt - 2; i >= 1; i--) { fr t; FR_MUL(t, z, q[i]); FR_ADD(q[i - 1], co[i], t); }  /* src/polynomial.rs:168-179 */
    fr rebuilt; FR_MUL(rebuilt, z, q[0]); FR_SUB(rebuilt, FR_ZERO, rebuilt);
    if (!eq(rebuilt.v, co[0].v, NR)) rc = 3;
    else if (nt - 1 > s->n) rc = 1;
    else commit_fr(s, q, nt - 1, threads, &acc);
    free(q);
  }
  if (rc == 0) p1_compress(out, &acc);
  free(co);
  return rc;
}
