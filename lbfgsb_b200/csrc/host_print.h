// The reference's text output, produced on the host from the mirrored state block:
// prn1lb (src/lbfgsb.f90:2363-2412), prn2lb (:2432-2468), prn3lb (:2487-2579), the iterate-0 lines of
// mainlb (:584-589), the messages of active (:1031-1038) and the restart / skip / backtrack messages of
// mainlb, cauchy, subsm and lnsrlb, which the scalar kernels log in order (common.cuh: EV_*).
//
// Levels (test/driver1.f90:140-150): iprint < 0 nothing; 0 one summary at the end; 0 < iprint < 99
// f and |proj g| every iprint iterations + the file iterate.dat; 99 every iteration; 100 also the final x;
// > 100 also x and g at every iteration.  The per-segment trace that the reference's sequential Cauchy
// loop prints at iprint >= 99 (:1367-1528, :2017-2056) has no counterpart in the sorted walk and is not
// produced.  Fortran edit descriptors are reproduced exactly (1P,Dw.d / 1P,Ew.d / Iw with asterisks on
// overflow); list-directed items (`write(*,*)`) are compiler dependent and follow the layout of the
// reference's golden files (test/OUTPUTS/output_90_1): integers in 12 columns, reals as ES24.15E3.
// The phase timers (cachyt, sbtime, lnscht) are reported as zero; `Total User time` is the wall time
// since START.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <string>

namespace lbprint {

// 1P,<letter>w.d  (letter 'D' or 'E'); two-digit exponent, the letter is dropped for three digits
inline std::string fmt_1p(double v, int w, int d, char letter) {
    std::string body;
    if (std::isnan(v)) body = "NaN";
    else if (std::isinf(v)) body = v > 0 ? "Infinity" : "-Infinity";
    else {
        char buf[64];
        snprintf(buf, sizeof buf, "%.*E", d, v);
        std::string s(buf);
        const size_t e = s.find('E');
        std::string mant = s.substr(0, e);
        int ex = atoi(s.c_str() + e + 1);
        char eb[16];
        if (ex > -100 && ex < 100) snprintf(eb, sizeof eb, "%c%c%02d", letter, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        else snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        body = mant + eb;
    }
    if ((int)body.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)w - body.size(), ' ') + body;
}
inline std::string fmt_i(long long v, int w) {
    char buf[32];
    snprintf(buf, sizeof buf, "%lld", v);
    std::string s(buf);
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)w - s.size(), ' ') + s;
}
inline std::string list_int(long long v) { return fmt_i(v, 12); }
inline std::string list_real(double v) {
    if (std::isnan(v)) return "                     NaN";
    if (std::isinf(v)) return v > 0 ? "                Infinity" : "               -Infinity";
    char buf[64];
    snprintf(buf, sizeof buf, "%.15E", v);
    std::string s(buf);
    const size_t e = s.find('E');
    int ex = atoi(s.c_str() + e + 1);
    char eb[16];
    snprintf(eb, sizeof eb, "E%c%03d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    std::string body = s.substr(0, e) + eb;
    return (body.size() < 24 ? std::string(24 - body.size(), ' ') : std::string(" ")) + body;
}

struct Ctx {
    int iprint = -1;
    FILE* itf = nullptr;
    std::string itname = "iterate.dat";
    std::chrono::steady_clock::time_point t0;
    char word[4] = {'-', '-', '-', 0};
    ~Ctx() { if (itf) fclose(itf); }
    void open_file() {   // :481-489
        if (itf) { fclose(itf); itf = nullptr; }
        if (iprint >= 1) itf = fopen(itname.c_str(), "w");
    }
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// '(/,a4,1p,6(1x,d11.4),/,(4x,1p,6(1x,d11.4)))'
template <typename T>
inline void print_vec(FILE* f, const char* tag4, const T* v, long long n) {
    fprintf(f, "\n%4s", tag4);
    for (long long i = 0; i < n; ++i) {
        if (i > 0 && i % 6 == 0) fprintf(f, "\n    ");
        fprintf(f, " %s", fmt_1p((double)v[i], 11, 4, 'D').c_str());
    }
    fprintf(f, "\n");
}

// prn1lb (:2363-2412) + the messages of active (:1031-1038)
template <typename T>
inline void prn1lb(Ctx& c, long long n, int m, double epsmch, bool prjctd, bool cnstnd, long long nbdd,
                   const T* l, const T* u, const T* x) {
    if (c.iprint >= 0) {
        printf("RUNNING THE L-BFGS-B CODE\n\n           * * *\n\nMachine precision =%s\n", fmt_1p(epsmch, 10, 3, 'D').c_str());
        printf(" N = %s     M = %s\n", list_int(n).c_str(), list_int(m).c_str());
        if (c.iprint >= 1 && c.itf) {
            fprintf(c.itf,
                    "RUNNING THE L-BFGS-B CODE\n\nit    = iteration number\nnf    = number of function evaluations\n"
                    "nseg  = number of segments explored during the Cauchy search\n"
                    "nact  = number of active bounds at the generalized Cauchy point\n"
                    "sub   = manner in which the subspace minimization terminated:\n"
                    "        con = converged, bnd = a bound was reached\n"
                    "itls  = number of iterations performed in the line search\nstepl = step length used\n"
                    "tstep = norm of the displacement (total step)\nprojg = norm of the projected gradient\n"
                    "f     = function value\n\n           * * *\n\nMachine precision =%s\n",
                    fmt_1p(epsmch, 10, 3, 'D').c_str());
            fprintf(c.itf, " N = %s     M = %s\n", list_int(n).c_str(), list_int(m).c_str());
            fprintf(c.itf, "\n   it   nf  nseg  nact  sub  itls  stepl    tstep     projg        f\n");
        }
        if (c.iprint > 100 && l && u && x) {
            print_vec<T>(stdout, "L =", l, n);
            print_vec<T>(stdout, "X0 =", x, n);
            print_vec<T>(stdout, "U =", u, n);
        }
    }
    if (c.iprint >= 0) {
        if (prjctd) printf(" The initial X is infeasible.  Restart with its projection.\n");
        if (!cnstnd) printf(" This problem is unconstrained.\n");
    }
    if (c.iprint > 0) printf("\nAt X0 %s variables are exactly at the bounds\n", fmt_i(nbdd, 9).c_str());
}

// mainlb :584-589
inline void iterate0(Ctx& c, int iter, int nfgv, double f, double sbgnrm) {
    if (c.iprint < 1) return;
    printf("\nAt iterate%s    f= %s    |proj g|= %s\n", fmt_i(iter, 5).c_str(), fmt_1p(f, 12, 5, 'D').c_str(),
           fmt_1p(sbgnrm, 12, 5, 'D').c_str());
    if (c.itf)
        fprintf(c.itf, " %s %s     -     -   -     -     -        -    %s %s\n", fmt_i(iter, 4).c_str(), fmt_i(nfgv, 4).c_str(),
                fmt_1p(sbgnrm, 10, 3, 'D').c_str(), fmt_1p(f, 10, 3, 'D').c_str());
}

// one logged message (common.cuh EV_*)
inline void event(Ctx& c, int code, double a, double b) {
    const int ip = c.iprint;
    static const char* refresh = "   refresh the lbfgs memory and restart the iteration.";
    switch (code) {
        case 1: if (ip >= 99) printf("\n\nITERATION %s\n", fmt_i((long long)a, 5).c_str()); break;
        case 2: if (ip >= 0) printf(" Subgnorm = 0.  GCP = X.\n"); break;
        case 3: case 5: if (ip >= 1) printf("\n Singular triangular system detected;\n%s\n", refresh); break;
        case 4: if (ip >= 1) printf("\n Nonpositive definiteness in Cholesky factorization in formk;\n%s\n", refresh); break;
        case 6: if (ip >= 0) printf(" Positive dir derivative in projection \n Using the backtracking step \n"); break;
        case 7: printf("  ascent direction in projection gd = %s\n", list_real(a).c_str()); break;   // unconditional (:2250)
        case 8: if (ip >= 1) printf("\n Bad direction in the line search;\n%s\n", refresh); break;
        case 9: if (ip >= 1) printf("  ys=%s  -gs=%s BFGS update SKIPPED\n", fmt_1p(a, 10, 3, 'E').c_str(), fmt_1p(b, 10, 3, 'E').c_str()); break;
        case 10: if (ip >= 1) printf("\n Nonpositive definiteness in Cholesky factorization in formt;\n%s\n", refresh); break;
        default: break;
    }
}

// prn2lb (:2432-2468)
template <typename T>
inline void prn2lb(Ctx& c, long long n, const T* x, const T* g, double f, int iter, int nfgv, long long nact, double sbgnrm,
                   long long nseg, int iword, int iback, double stp, double xstep) {
    const char* w = (iword == 0) ? "con" : (iword == 1) ? "bnd" : (iword == 5) ? "TNT" : "---";
    snprintf(c.word, sizeof c.word, "%s", w);
    const int ip = c.iprint;
    if (ip >= 99) {
        printf(" LINE SEARCH%s  times; norm of step = %s\n", list_int(iback).c_str(), list_real(xstep).c_str());
        printf("\nAt iterate%s    f= %s    |proj g|= %s\n", fmt_i(iter, 5).c_str(), fmt_1p(f, 12, 5, 'D').c_str(),
               fmt_1p(sbgnrm, 12, 5, 'D').c_str());
        if (ip > 100 && x && g) { print_vec<T>(stdout, "X =", x, n); print_vec<T>(stdout, "G =", g, n); }
    } else if (ip > 0) {
        if (iter % ip == 0)
            printf("\nAt iterate%s    f= %s    |proj g|= %s\n", fmt_i(iter, 5).c_str(), fmt_1p(f, 12, 5, 'D').c_str(),
                   fmt_1p(sbgnrm, 12, 5, 'D').c_str());
    }
    if (ip >= 1 && c.itf)
        fprintf(c.itf, " %s %s %s %s  %3s %s  %s  %s %s %s\n", fmt_i(iter, 4).c_str(), fmt_i(nfgv, 4).c_str(),
                fmt_i(nseg, 5).c_str(), fmt_i(nact, 5).c_str(), w, fmt_i(iback, 4).c_str(), fmt_1p(stp, 7, 1, 'D').c_str(),
                fmt_1p(xstep, 7, 1, 'D').c_str(), fmt_1p(sbgnrm, 10, 3, 'D').c_str(), fmt_1p(f, 10, 3, 'D').c_str());
}

inline void info_text(FILE* f, int info) {
    switch (info) {
        case -1: fprintf(f, "\n Matrix in 1st Cholesky factorization in formk is not Pos. Def.\n"); break;
        case -2: fprintf(f, "\n Matrix in 2st Cholesky factorization in formk is not Pos. Def.\n"); break;
        case -3: fprintf(f, "\n Matrix in the Cholesky factorization in formt is not Pos. Def.\n"); break;
        case -4: fprintf(f, "\n Derivative >= 0, backtracking line search impossible.\n   Previous x, f and g restored.\n"
                            " Possible causes: 1 error in function or gradient evaluation;\n"
                            "                  2 rounding errors dominate computation.\n"); break;
        case -5: fprintf(f, "\n Warning:  more than 10 function and gradient\n   evaluations in the last line search.  Termination\n"
                            "   may possibly be caused by a bad search direction.\n"); break;
        case -8: fprintf(f, "\n The triangular system is singular.\n"); break;
        case -9: fprintf(f, "\n Line search cannot locate an adequate point after 20 function\n"
                            "  and gradient evaluations.  Previous x, f and g restored.\n"
                            " Possible causes: 1 error in function or gradient evaluation;\n"
                            "                  2 rounding error dominate computation.\n"); break;
        default: break;
    }
}

// prn3lb (:2487-2579)
template <typename T>
inline void prn3lb(Ctx& c, long long n, const T* x, double f, const char* task60, int info, int iter, int nfgv, long long nintol,
                   int nskip, long long nact, double sbgnrm, double time, long long nseg, int iback, double stp, double xstep,
                   long long k) {
    const int ip = c.iprint;
    const bool err = memcmp(task60, "ERROR", 5) == 0;
    if (!err && ip >= 0) {
        printf("\n           * * *\n\nTit   = total number of iterations\nTnf   = total number of function evaluations\n"
               "Tnint = total number of segments explored during Cauchy searches\nSkip  = number of BFGS updates skipped\n"
               "Nact  = number of active bounds at final generalized Cauchy point\n"
               "Projg = norm of the final projected gradient\nF     = final function value\n\n           * * *\n");
        printf("\n   N    Tit     Tnf  Tnint  Skip  Nact     Projg        F\n");
        printf("%s %s %s %s  %s %s  %s  %s\n", fmt_i(n, 5).c_str(), fmt_i(iter, 6).c_str(), fmt_i(nfgv, 6).c_str(),
               fmt_i(nintol, 6).c_str(), fmt_i(nskip, 4).c_str(), fmt_i(nact, 5).c_str(), fmt_1p(sbgnrm, 10, 3, 'D').c_str(),
               fmt_1p(f, 10, 3, 'D').c_str());
        if (ip >= 100 && x) print_vec<T>(stdout, "X =", x, n);
        if (ip >= 1) printf("  F =%s\n", list_real(f).c_str());
    }
    if (ip >= 0) {
        printf("\n%.60s\n", task60);
        if (info == -6) printf("  Input nbd(%s ) is invalid.\n", list_int(k).c_str());
        else if (info == -7) printf("  l(%s ) > u(%s ).  No feasible solution.\n", list_int(k).c_str(), list_int(k).c_str());
        else info_text(stdout, info);
        if (ip >= 1)
            printf("\n Cauchy                time%s seconds.\n Subspace minimization time%s seconds.\n Line search           time%s seconds.\n",
                   fmt_1p(0.0, 10, 3, 'E').c_str(), fmt_1p(0.0, 10, 3, 'E').c_str(), fmt_1p(0.0, 10, 3, 'E').c_str());
        printf("\n Total User time%s seconds.\n\n", fmt_1p(time, 10, 3, 'E').c_str());
        if (ip >= 1 && c.itf) {
            if (info == -4 || info == -9)
                fprintf(c.itf, " %s %s %s %s  %3s %s  %s  %s      -          -\n", fmt_i(iter, 4).c_str(), fmt_i(nfgv, 4).c_str(),
                        fmt_i(nseg, 5).c_str(), fmt_i(nact, 5).c_str(), c.word, fmt_i(iback, 4).c_str(),
                        fmt_1p(stp, 7, 1, 'D').c_str(), fmt_1p(xstep, 7, 1, 'D').c_str());
            fprintf(c.itf, "\n%.60s\n", task60);
            if (info == -4) info_text(stdout, info);   // the reference writes this one to stdout (:2546)
            else if (info != -6 && info != -7) info_text(c.itf, info);
            fprintf(c.itf, "\n Total User time%s seconds.\n\n", fmt_1p(time, 10, 3, 'E').c_str());
            fflush(c.itf);
        }
    }
    fflush(stdout);
}

}  // namespace lbprint
