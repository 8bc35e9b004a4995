// CPU emulation of assemble_sumfact_hex_p2_poisson_kernel (neopz_b200/csrc/sumfact_hex.cuh): the kernel's phase functions are
// run with loops over the thread index instead of threads, barriers become the loop boundaries.  Test infrastructure
// (tests/test_sumfact_emulation.py compares the result with the reference's element matrices); not part of the product.
//   g++ -O2 -shared -fPIC tools/sumfact_emu.cpp -o tools/bin/libsumfact_emu.so
#include <cmath>
#include <cstring>
#define SF_HD inline
struct double2 { double x, y; };
#include "../neopz_b200/csrc/sumfact_hex.cuh"

double h_sfF3[4 * 9 * 3];

// X[8][3], dng[27][3][8], qw[27], x1d[3] (the line rule), scale = fScale.  K[27][27] (both triangles filled), ef_w[27] = w|detJ|.
// Returns the number of distinct upper-triangle entries written (378 when the work-item map is complete and free of duplicates),
// negative when an entry was written twice.
// mode: unused (kept for the binding)
extern "C" int sf_emulate_mode(const double *X, const double *dng, const double *qw, const double *x1d, double scale, double *K, double *Wd, int mode);
extern "C" int sf_emulate(const double *X, const double *dng, const double *qw, const double *x1d, double scale, double *K, double *Wd) {
    return sf_emulate_mode(X, dng, qw, x1d, scale, K, Wd, 0);
}
extern "C" int sf_emulate_mode(const double *X, const double *dng, const double *qw, const double *x1d, double scale, double *K, double *Wd, int mode) {
    double aux[sf::AUX_LEN];
    sf::build_tables(x1d, aux, h_sfF3);
    double Msm[6 * 27], S1[2][sf::NITEM + 2];
    for (int t = 0; t < 27; t++) sf::geometry(t, X, dng + (size_t)t * 24, 1, qw[t], scale, Msm, Wd);
    double acc[sf::NTHREADS][9];
    std::memset(acc, 0, sizeof(acc));
    for (int c = 0; c < 9; c++) {
        const int e = c / 3, f = c % 3;
        for (int t = 0; t < sf::NITEM; t++) sf::stage1(t, e, f, aux + sf::AUX_F1 + (sf::variant(0, e, f) * 6 + t / 9) * 3, Msm, S1[c & 1]);
        for (int t = 0; t < sf::NITEM; t++) sf::stage23(t, e, f, aux + sf::AUX_F2 + (sf::variant(1, e, f) * 9 + t % 9) * 3, S1[c & 1], acc[t]);
    }
    int written[27][27];
    std::memset(written, 0, sizeof(written));
    int count = 0;
    for (int t = 0; t < sf::NTHREADS; t++)
        for (int k = 0; k < 9; k++) {
            int i, j;
            if (!sf::entry_of(t, k, i, j)) continue;
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            if (written[lo][hi]++) return -1 - (lo * 27 + hi);
            K[i * 27 + j] = acc[t][k];
            K[j * 27 + i] = acc[t][k];
            count++;
        }
    return count;
}
