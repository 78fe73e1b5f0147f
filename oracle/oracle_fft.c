/*
 * oracle_fft.c -- TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
 *
 * CPU restatement, in plain C, of the slab-decomposed 3-D C2C FFT of the reference's
 * `3dmpifft_opt` hot path.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  P "devices" are simulated inside one process: each
 * device p owns two buffers buf1[p], buf2[p] exactly like the plan's bufferDev1/bufferDev2
 * (3dmpifft_opt/include/fft_mpi_3d_api.h:24), and the four stages t0..t3 are executed
 * device-by-device with the same index maps, counts and offsets as the reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm
 * may use this file.
 *
 * Parity pin: the reference's own tests pin only the fwd+bwd round trip
 * (3dmpifft_opt/fftSpeed3d_c2c.cpp:79-91).  The 1-D engine below is pinned against the
 * golden vectors the reference tree carries (heffte/heffteBenchmark/test/test_units_nompi.cpp:92-190,
 * test_units_stock.cpp:229-255) in tests/test_oracle.py, and against numpy pocketfft.  The stage functions are pinned on
 * EXECUTED reference code: tests/test_oracle_ref3d.py compares both buffers of every device after every stage with the
 * reference's own fft_mpi_3d_api.cpp + kernel_func.cpp + cuTranspose kernels run on the CPU (oracle/ref_3dmpifft), and
 * tests/test_oracle_ref.py compares whole spectra with the reference tree's heFFTe (oracle/ref_heffte).
 *
 * Math convention (templateFFT/src/templateFFT.cpp:5121-5141 LUT holds e^{+i theta};
 * :338 forward kernels conjugate it; :5946 normalize=0):
 *   forward  X[k] = sum_j x[j] e^{-2 pi i jk/N},   backward X[k] = sum_j x[j] e^{+2 pi i jk/N},
 *   both unnormalised.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;
typedef long long i64;

#define ORACLE_FORWARD 1   /* 3dmpifft_opt/include/fft_mpi_common.h:18 */
#define ORACLE_BACKWARD (-1) /* fft_mpi_common.h:19 */

/* ------------------------------------------------------------------------------------------
 * Radix schedule.  templateFFT.cpp:3956-3963 factors N over {2..13}; :4540-4550 merges three
 * 2s into an 8, then two 2s into a 4; :4580-4588 orders stages by descending radix.
 * Returns the number of stages, 0 if N has a prime factor > 13 (the reference returns
 * FFT_ERROR_UNSUPPORTED_RADIX there, templateFFT.cpp:3964).
 * Checked against the EXECUTED generator (tests/test_oracle_ref3d.py, oracle/ref_3dmpifft): exact for powers of 2, 3, 5 and 7
 * (512 = 8.8.8, 1024 = 8.8.8.2, ...); for mixed lengths the generator applies the :4540-4550 merge only when the per-thread
 * register count of the other radix is a multiple of 8 / 4 (768 = 4.4.4.4.3 there, 8.8.4.3 here) and it has no radix 11 / 13
 * kernels.  The schedule changes the rounding, not the transform; the values are compared with the generated kernels' outputs.
 * ---------------------------------------------------------------------------------------- */
int oracle_radix_schedule(int n, int *radix)
{
    int mult[14];
    memset(mult, 0, sizeof(mult));
    int t = n;
    for (int i = 2; i < 14; i++)
        while (t % i == 0) { t /= i; mult[i]++; }
    if (t != 1) return 0;
    /* composite entries 4,6,8,9,10,12 can never be hit above because 2 and 3 are divided out
       first; the merge below is what creates 8s and 4s */
    mult[8] = mult[2] / 3; mult[2] -= 3 * mult[8];
    mult[4] = mult[2] / 2; mult[2] -= 2 * mult[4];
    int ns = 0;
    for (int i = 13; i > 1; i--)
        for (int k = 0; k < mult[i]; k++) radix[ns++] = i;
    return ns;
}

/* exact-ish twiddle: e^{sign * 2 pi i k / n}, evaluated in long double then rounded once */
static inline cplx twiddle(i64 k, i64 n, int sign)
{
    long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)(k % n) / (long double)n;
    cplx w; w.re = (double)cosl(a); w.im = (double)(sign * sinl(a));
    return w;
}

static inline cplx cmul(cplx a, cplx b)
{
    cplx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r;
}

/* One Stockham autosort stage of radix r over a contiguous line of n points
 * (structure of templateFFT's appendRadixStage/appendRadixShuffle pair, templateFFT.cpp:1871, 2466):
 * butterfly j reads in[j + m*n/r], applies the stage twiddle w^(k*m), k = j mod ns,
 * does the r-point DFT (inlineRadixKernelFFT, templateFFT.cpp:315-1075) and writes
 * out[(j-k)*r + k + m*ns].  `root` holds e^{sign 2 pi i q / r} for the small DFT,
 * `tw` the n-point table e^{sign 2 pi i q / n}. */
static void stockham_stage(const cplx *in, cplx *out, int n, int r, int ns,
                           const cplx *stw, const cplx *root)
{
    /* stw[k*(r-1) + (m-1)] = e^{sign 2 pi i k m / (ns r)}: the per-stage LUT of the reference
       (templateFFT.cpp:5121-5141), so the inner loop has no index arithmetic beyond k */
    const int nb = n / r;            /* butterflies */
    cplx v[13], y[13];
    int k = 0;
    for (int j = 0; j < nb; j++, k = (k + 1 == ns) ? 0 : k + 1) {
        v[0] = in[j];
        if (k == 0) for (int m = 1; m < r; m++) v[m] = in[j + m * nb];
        else { const cplx *w = stw + (size_t)k * (r - 1); for (int m = 1; m < r; m++) v[m] = cmul(in[j + m * nb], w[m - 1]); }
        if (r == 2) {
            y[0].re = v[0].re + v[1].re; y[0].im = v[0].im + v[1].im;
            y[1].re = v[0].re - v[1].re; y[1].im = v[0].im - v[1].im;
        } else if (r == 4) {
            /* root[1] = e^{sign i pi/2} = (0, sign) */
            cplx a = { v[0].re + v[2].re, v[0].im + v[2].im };
            cplx b = { v[0].re - v[2].re, v[0].im - v[2].im };
            cplx c = { v[1].re + v[3].re, v[1].im + v[3].im };
            cplx d = { v[1].re - v[3].re, v[1].im - v[3].im };
            cplx dj = cmul(d, root[1]);
            y[0].re = a.re + c.re; y[0].im = a.im + c.im;
            y[2].re = a.re - c.re; y[2].im = a.im - c.im;
            y[1].re = b.re + dj.re; y[1].im = b.im + dj.im;
            y[3].re = b.re - dj.re; y[3].im = b.im - dj.im;
        } else if (r == 8) {
            /* decimation in frequency: even outputs = DFT4(v[n]+v[n+4]), odd = DFT4((v[n]-v[n+4]) w8^n) */
            cplx a[4], b[4];
            for (int q = 0; q < 4; q++) {
                a[q].re = v[q].re + v[q + 4].re; a[q].im = v[q].im + v[q + 4].im;
                cplx d = { v[q].re - v[q + 4].re, v[q].im - v[q + 4].im };
                b[q] = q ? cmul(d, root[q]) : d;
            }
            for (int h = 0; h < 2; h++) {
                const cplx *u = h ? b : a;
                cplx s0 = { u[0].re + u[2].re, u[0].im + u[2].im };
                cplx s1 = { u[0].re - u[2].re, u[0].im - u[2].im };
                cplx s2 = { u[1].re + u[3].re, u[1].im + u[3].im };
                cplx s3 = { u[1].re - u[3].re, u[1].im - u[3].im };
                cplx s3j = cmul(s3, root[2]);
                y[h + 0].re = s0.re + s2.re; y[h + 0].im = s0.im + s2.im;
                y[h + 4].re = s0.re - s2.re; y[h + 4].im = s0.im - s2.im;
                y[h + 2].re = s1.re + s3j.re; y[h + 2].im = s1.im + s3j.im;
                y[h + 6].re = s1.re - s3j.re; y[h + 6].im = s1.im - s3j.im;
            }
        } else {
            for (int q = 0; q < r; q++) {
                cplx s = v[0];
                for (int m = 1; m < r; m++) {
                    cplx p = cmul(v[m], root[(q * m) % r]);
                    s.re += p.re; s.im += p.im;
                }
                y[q] = s;
            }
        }
        const int j0 = (j - k) * r + k;
        for (int m = 0; m < r; m++) out[j0 + m * ns] = y[m];
    }
}

typedef struct {
    int n, nstages, radix[32], sign;
    cplx *tw;          /* n entries (plain-DFT fallback only) */
    cplx *root[32];    /* per stage r entries */
    cplx *stw[32];     /* per stage ns*(r-1) twiddles */
} fft_plan1d;

static int plan1d_init(fft_plan1d *p, int n, int sign)
{
    p->n = n; p->sign = sign;
    p->nstages = oracle_radix_schedule(n, p->radix);
    p->tw = (cplx *)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; k++) p->tw[k] = twiddle(k, n, sign);
    int ns = 1;
    for (int s = 0; s < p->nstages; s++) {
        int r = p->radix[s];
        p->root[s] = (cplx *)malloc(sizeof(cplx) * (size_t)r);
        for (int q = 0; q < r; q++) p->root[s][q] = twiddle(q, r, sign);
        p->stw[s] = (cplx *)malloc(sizeof(cplx) * (size_t)ns * (size_t)(r - 1));
        for (int k = 0; k < ns; k++)
            for (int m = 1; m < r; m++) p->stw[s][(size_t)k * (r - 1) + (m - 1)] = twiddle((i64)k * m, (i64)ns * r, sign);
        ns *= r;
    }
    return p->nstages;
}

static void plan1d_free(fft_plan1d *p)
{
    free(p->tw);
    for (int s = 0; s < p->nstages; s++) { free(p->root[s]); free(p->stw[s]); }
}

/* transform one contiguous line in `a` (result returned in `a`), `b` is scratch of n */
static void fft_line(const fft_plan1d *p, cplx *a, cplx *b)
{
    const int n = p->n;
    if (n == 1) return;
    if (p->nstages == 0) { /* unsupported radix in the reference; plain DFT keeps the oracle total */
        for (int k = 0; k < n; k++) {
            cplx s = { 0, 0 };
            for (int j = 0; j < n; j++) { cplx t = cmul(a[j], p->tw[((i64)j * k) % n]); s.re += t.re; s.im += t.im; }
            b[k] = s;
        }
        memcpy(a, b, sizeof(cplx) * (size_t)n);
        return;
    }
    cplx *src = a, *dst = b;
    int ns = 1;
    for (int s = 0; s < p->nstages; s++) {
        stockham_stage(src, dst, n, p->radix[s], ns, p->stw[s], p->root[s]);
        ns *= p->radix[s];
        cplx *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, sizeof(cplx) * (size_t)n);
}

/* Batched 1-D transform: `nlines` lines of `n` points; line l starts at
 * base + (l / inner) * outer_dist + (l % inner) * inner_dist, points are `stride` apart.
 * sign = -1 forward (e^{-i..}), +1 backward. Exposed for the golden-vector tests. */
void oracle_fft_batch(cplx *base, int n, i64 stride, i64 nlines, i64 inner, i64 inner_dist,
                      i64 outer_dist, int sign)
{
    fft_plan1d plan;
    plan1d_init(&plan, n, sign);
    enum { BLK = 8 };   /* columns gathered together: 8 complex doubles = two cache lines per row */
#pragma omp parallel
    {
        cplx *a = (cplx *)malloc(sizeof(cplx) * (size_t)n * BLK);
        cplx *b = (cplx *)malloc(sizeof(cplx) * (size_t)n);
        if (stride == 1) {
#pragma omp for schedule(static)
            for (i64 l = 0; l < nlines; l++)
                fft_line(&plan, base + (l / inner) * outer_dist + (l % inner) * inner_dist, b);
        } else if (inner_dist == 1) {
            const i64 nblk = (inner + BLK - 1) / BLK, groups = nlines / inner;
#pragma omp for schedule(static)
            for (i64 w = 0; w < groups * nblk; w++) {
                const i64 gI = w / nblk, c0 = (w % nblk) * BLK;
                const int nc = (int)((inner - c0) < BLK ? (inner - c0) : BLK);
                cplx *p0 = base + gI * outer_dist + c0;
                for (int i = 0; i < n; i++)
                    for (int c = 0; c < nc; c++) a[(size_t)c * n + i] = p0[(i64)i * stride + c];
                for (int c = 0; c < nc; c++) fft_line(&plan, a + (size_t)c * n, b);
                for (int i = 0; i < n; i++)
                    for (int c = 0; c < nc; c++) p0[(i64)i * stride + c] = a[(size_t)c * n + i];
            }
        } else {
#pragma omp for schedule(static)
            for (i64 l = 0; l < nlines; l++) {
                cplx *p0 = base + (l / inner) * outer_dist + (l % inner) * inner_dist;
                for (int i = 0; i < n; i++) a[i] = p0[(i64)i * stride];
                fft_line(&plan, a, b);
                for (int i = 0; i < n; i++) p0[(i64)i * stride] = a[i];
            }
        }
        free(a); free(b);
    }
    plan1d_free(&plan);
}

/* ------------------------------------------------------------------------------------------
 * Slab bookkeeping.
 * ---------------------------------------------------------------------------------------- */
static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

/* getProperDeviceNum, fft_mpi_3d_api.cpp:232-272 (single node: mpiSize = 1): shrink the device
 * count so that ceil-blocks of N0 leave no empty device. */
int oracle_proper_device_num(i64 n0, int wanted)
{
    int dev = wanted;
    if (n0 % wanted != 0) {
        i64 per = n0 / wanted + 1;
        dev = (int)(n0 / per);
        if (n0 % per != 0) dev += 1;
    }
    return dev;
}

/* getDataCountForNode, fft_mpi_3d_api.cpp:274-287 */
void oracle_data_counts(const i64 N[3], int P, i64 *count)
{
    i64 normal = ceil_div(N[0], P) * N[1] * N[2];
    for (int i = 0; i < P; i++)
        count[i] = (i == P - 1) ? N[0] * N[1] * N[2] - normal * (P - 1) : normal;
}

/* getMaxDataCount, fft_mpi_3d_api.cpp:289-316 */
i64 oracle_max_data_count(i64 n0, i64 n1, i64 n2, int P, int is_last)
{
    i64 n0d, n1d;
    if (is_last) { n0d = n0 - (P - 1) * ceil_div(n0, P); n1d = n1 - (P - 1) * ceil_div(n1, P); }
    else { n0d = ceil_div(n0, P); n1d = ceil_div(n1, P); }
    i64 a = n0d * n1 * n2, b = n0 * n1d * n2;
    return a <= b ? b : a;
}

/* exchange table of one device, fft_mpi_3d_api.cpp:84-133. arrays of P entries. */
void oracle_exchange_table(i64 n0, i64 n1, i64 n2, int P, int dev, int direction,
                           i64 *scount, i64 *soffset, i64 *rcount, i64 *roffset)
{
    const i64 xd = ceil_div(n0, P), yd = ceil_div(n1, P);
    const i64 lastN0 = n0 - (P - 1) * xd, lastN1 = n1 - (P - 1) * yd;
    const int last = P - 1;
    if (dev != last) {
        i64 normal = xd * yd * n2;
        for (int i = 0; i < P - 1; i++) { rcount[i] = scount[i] = normal; roffset[i] = soffset[i] = i * normal; }
        if (direction == ORACLE_FORWARD) { rcount[last] = lastN0 * yd * n2; scount[last] = xd * lastN1 * n2; }
        else { rcount[last] = xd * lastN1 * n2; scount[last] = lastN0 * yd * n2; }
        roffset[last] = soffset[last] = (i64)last * normal;
    } else {
        i64 nr, nsnd;
        if (direction == ORACLE_FORWARD) { nr = xd * lastN1 * n2; nsnd = lastN0 * yd * n2; }
        else { nr = lastN0 * yd * n2; nsnd = xd * lastN1 * n2; }
        for (int i = 0; i < P - 1; i++) { rcount[i] = nr; scount[i] = nsnd; roffset[i] = i * nr; soffset[i] = i * nsnd; }
        rcount[last] = scount[last] = lastN0 * lastN1 * n2;
        roffset[last] = (i64)last * nr; soffset[last] = (i64)last * nsnd;
    }
}

/* ------------------------------------------------------------------------------------------
 * The four stages, one device at a time.
 * ---------------------------------------------------------------------------------------- */

/* t0: fftZY, fft_mpi_3d_api.cpp:466-522 -- n0_l independent in-place 2-D transforms of N1 x N2
 * (axis-0 = Z contiguous, then axis-1 = Y strided by N2; templateFFT.cpp:6212-6233). */
void oracle_stage_fftZY(cplx *buf, i64 n0l, i64 n1, i64 n2, int direction)
{
    const int sign = direction == ORACLE_FORWARD ? -1 : +1;
    /* Z: n0l*n1 contiguous lines */
    oracle_fft_batch(buf, (int)n2, 1, n0l * n1, n0l * n1, n2, 0, sign);
    /* Y: for each plane, n2 lines with stride n2 */
    oracle_fft_batch(buf, (int)n1, n2, n0l * n2, n2, 1, n1 * n2, sign);
}

/* t1: localTransposeUneven -> slab_local_transpose_z_to_x_uneven_{forward,backward}_optimized,
 * fft_mpi_3d_api.cpp:575-608, kernel_func.cpp:73-100.
 * forward: out[packed] = in[natural]; backward: out[natural] = in[packed]. */
void oracle_stage_pack(const cplx *in, cplx *out, i64 x_size, i64 n1, i64 n2, int P, int direction)
{
    const i64 yd = ceil_div(n1, P), y_last = n1 - (P - 1) * yd;
#pragma omp parallel for schedule(static)
    for (i64 x = 0; x < x_size; x++)
        for (i64 y = 0; y < n1; y++) {
            const i64 q = y / yd;
            const i64 w = (q == P - 1) ? y_last : yd;
            const i64 nat = (x * n1 + y) * n2;
            const i64 pk = x_size * yd * n2 * q + x * w * n2 + (y % yd) * n2;
            if (direction == ORACLE_FORWARD) memcpy(out + pk, in + nat, sizeof(cplx) * (size_t)n2);
            else memcpy(out + nat, in + pk, sizeof(cplx) * (size_t)n2);
        }
}

/* t2: slabAlltoall, fft_mpi_3d_api.cpp:610-672. Sender s pushes chunk i of its buf2
 * (offset soffset[i], scount[i] elements) into device i's buf1 at the receive offset of
 * lines 618-625 (which uses the *sender's* global index). */
void oracle_stage_alltoall(cplx **buf2, cplx **buf1, i64 n0, i64 n1, i64 n2, int P, int direction)
{
    const i64 xd = ceil_div(n0, P), yd = ceil_div(n1, P);
    const i64 lastN0 = n0 - (P - 1) * xd, lastN1 = n1 - (P - 1) * yd;
    i64 *sc = (i64 *)malloc(sizeof(i64) * 4 * (size_t)P), *so = sc + P, *rc = so + P, *ro = rc + P;
    for (int s = 0; s < P; s++) {
        oracle_exchange_table(n0, n1, n2, P, s, direction, sc, so, rc, ro);
        for (int i = 0; i < P; i++) {
            i64 recv_off;
            if (i == P - 1) recv_off = (direction == ORACLE_FORWARD) ? s * xd * lastN1 * n2 : s * lastN0 * yd * n2;
            else recv_off = s * xd * yd * n2;
            /* the copy is split over the host threads (a device-to-device copy in the reference); same bytes moved */
            {
                cplx *dst = buf1[i] + recv_off;
                const cplx *src = buf2[s] + so[i];
                const i64 cnt = sc[i], blk = 1 << 16;
#pragma omp parallel for schedule(static)
                for (i64 b0 = 0; b0 < cnt; b0 += blk)
                    memcpy(dst + b0, src + b0, sizeof(cplx) * (size_t)(cnt - b0 < blk ? cnt - b0 : blk));
            }
        }
    }
    free(sc);
}

/* t3: fftX, fft_mpi_3d_api.cpp:524-573.
 * forward: cut_transpose3d perm {2,0,1} (fast_transpose/kernels_201.cpp:46-57 with np0=N2,
 *   np1=n1_l, np2=N0): buf2[(y*N2+z)*N0 + x] = buf1[(x*n1_l + y)*N2 + z]; then in-place
 *   length-N0 transforms of the n1_l*N2 contiguous lines of buf2.
 * backward: in-place inverse transforms of the lines of buf1, then perm {1,2,0}
 *   (kernels_120.cpp:45-57 called with sizes {N0, n1_l, N2}, api.cpp:563-565):
 *   buf2[x*(n1_l*N2) + r] = buf1[r*N0 + x], r = y*N2+z. */
void oracle_stage_fftX(cplx *buf1, cplx *buf2, i64 n0, i64 n1l, i64 n2, int direction)
{
    const i64 rows = n1l * n2;
    if (direction == ORACLE_FORWARD) {
        enum { TB = 16 };
#pragma omp parallel for schedule(static)
        for (i64 rb = 0; rb < rows; rb += TB)
            for (i64 xb = 0; xb < n0; xb += TB)
                for (i64 r = rb; r < rb + TB && r < rows; r++)
                    for (i64 x = xb; x < xb + TB && x < n0; x++) buf2[r * n0 + x] = buf1[x * rows + r];
        oracle_fft_batch(buf2, (int)n0, 1, rows, rows, n0, 0, -1);
    } else {
        oracle_fft_batch(buf1, (int)n0, 1, rows, rows, n0, 0, +1);
        enum { TB = 16 };
#pragma omp parallel for schedule(static)
        for (i64 xb = 0; xb < n0; xb += TB)
            for (i64 rb = 0; rb < rows; rb += TB)
                for (i64 x = xb; x < xb + TB && x < n0; x++)
                    for (i64 r = rb; r < rb + TB && r < rows; r++) buf2[x * rows + r] = buf1[r * n0 + x];
    }
}

/* fft_mpi_execute_dft_3d_c2c, fft_mpi_3d_api.cpp:181-214, for all P simulated devices.
 * On entry buf1[p] holds device p's input (forward: x-slab [x_l][y][z]; backward: the forward
 * output [y_l][z][x]).  On exit buf2[p] holds the result (forward: [y_l][z][x]; backward:
 * [x_l][y][z], unnormalised).  stop_after in 0..3 stops after stage t<stop_after> of the
 * *forward order* (t0,t1,t2,t3) or, for backward, after its 1st..4th executed stage; pass 3
 * (or anything >= 3) for the whole transform. */
int oracle_slab_execute(int P, i64 n0, i64 n1, i64 n2, cplx **buf1, cplx **buf2, int direction, int stop_after)
{
    const i64 xd = ceil_div(n0, P), yd = ceil_div(n1, P);
    const i64 lastN0 = n0 - (P - 1) * xd, lastN1 = n1 - (P - 1) * yd;
    if (lastN0 < 1 || lastN1 < 1) return -1;
    if (direction == ORACLE_FORWARD) {
        for (int p = 0; p < P; p++) oracle_stage_fftZY(buf1[p], p == P - 1 ? lastN0 : xd, n1, n2, direction);
        if (stop_after == 0) return 0;
        for (int p = 0; p < P; p++) oracle_stage_pack(buf1[p], buf2[p], p == P - 1 ? lastN0 : xd, n1, n2, P, direction);
        if (stop_after == 1) return 0;
        oracle_stage_alltoall(buf2, buf1, n0, n1, n2, P, direction);
        if (stop_after == 2) return 0;
        for (int q = 0; q < P; q++) oracle_stage_fftX(buf1[q], buf2[q], n0, q == P - 1 ? lastN1 : yd, n2, direction);
    } else {
        for (int q = 0; q < P; q++) oracle_stage_fftX(buf1[q], buf2[q], n0, q == P - 1 ? lastN1 : yd, n2, direction);
        if (stop_after == 0) return 0;
        oracle_stage_alltoall(buf2, buf1, n0, n1, n2, P, direction);
        if (stop_after == 1) return 0;
        for (int p = 0; p < P; p++) oracle_stage_pack(buf1[p], buf2[p], p == P - 1 ? lastN0 : xd, n1, n2, P, direction);
        if (stop_after == 2) return 0;
        for (int p = 0; p < P; p++) oracle_stage_fftZY(buf2[p], p == P - 1 ? lastN0 : xd, n1, n2, direction);
    }
    return 0;
}

/* driver input ramp, fftSpeed3d_c2c.cpp:61-63: re = im = global linear index */
void oracle_fill_ramp(cplx *dst, i64 start, i64 count)
{
    for (i64 j = 0; j < count; j++) dst[j].re = dst[j].im = (double)(start + j);
}

/* heFFTe test input, heffte/heffteBenchmark/test/test_fft3d.h:19-27: std::minstd_rand(4242)
 * (Park-Miller, a = 48271, m = 2^31-1) through uniform_real_distribution<double>(0,1), consumed
 * in world order, imaginary part 0.  libstdc++'s generate_canonical<double,53> draws two
 * 31-bit-range samples per double: (s1-1 + (s2-1)*R) / R^2 with R = m-1.  `state` carries the
 * LCG state between calls (seed it with 4242). */
void oracle_fill_minstd(cplx *dst, i64 count, unsigned long long *state)
{
    const unsigned long long a = 48271ULL, m = 2147483647ULL;
    const double R = 2147483646.0; /* urng.max() - urng.min() + 1, arithmetic in double like libstdc++ */
    unsigned long long s = *state;
    for (i64 j = 0; j < count; j++) {
        s = (s * a) % m; double lo = (double)(s - 1);
        s = (s * a) % m; double hi = (double)(s - 1);
        double v = (lo + hi * R) / (R * R);
        if (v >= 1.0) v = nextafter(1.0, 0.0);
        dst[j].re = v; dst[j].im = 0.0;
    }
    *state = s;
}

/* driver error metric, fftSpeed3d_c2c.cpp:84-91: max_j |in_j - out_j/N^3| / 1e7 ; also returns
 * the plain absolute max error (heFFTe's gate, heffte/heffteBenchmark/benchmarks/speed3d.h:138-144). */
double oracle_roundtrip_error(const cplx *in, const cplx *out, i64 count, double n3, double *abs_err)
{
    double mx = -1.0;
    for (i64 j = 0; j < count; j++) {
        double a = in[j].re - out[j].re / n3, b = in[j].im - out[j].im / n3;
        double e = sqrt(a * a + b * b);
        if (e > mx) mx = e;
    }
    if (abs_err) *abs_err = mx;
    return mx / 1e7;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
