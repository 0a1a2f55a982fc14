// Minimal stand-in for <Rcpp.h>: just enough declarations for `g++ -fsyntax-only` to type-check r/src/swne_nmf_cuda.cpp
// in a container without R (tests/test_host_cpu.py::test_rcpp_shim_type_checks_against_the_c_abi).  It proves that every
// call of the shim into include/swne_nmf.h has the right arity and argument types; it is NOT Rcpp and nothing links to it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

typedef struct SEXPREC *SEXP;
inline bool Rf_isS4(SEXP) { return false; }

namespace Rcpp {
struct Range { Range(int, int) {} };
struct SlotProxy;
template <typename T>
struct Vector {
    Vector() {}
    explicit Vector(int) {}
    Vector(SEXP) {}
    Vector(const SlotProxy &) {}
    T *begin() { return nullptr; }
    int size() const { return 0; }
    T &operator[](int) { static T v; return v; }
    Vector operator[](const Range &) { return *this; }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
struct NumericMatrix {
    NumericMatrix() {}
    NumericMatrix(int, int) {}
    NumericMatrix(SEXP) {}
    double *begin() { return nullptr; }
    int nrow() const { return 0; }
    int ncol() const { return 0; }
    double &operator()(int, int) { static double v; return v; }
};
template <typename T> T clone(const T &x) { return x; }
struct SlotProxy {};
struct S4 {
    S4(SEXP) {}
    SlotProxy slot(const char *) { return SlotProxy(); }
};
struct NamedArg {
    template <typename T> NamedArg operator=(const T &) { return *this; }
};
struct NamedPlaceholder { NamedArg operator[](const char *) const { return NamedArg(); } };
static const NamedPlaceholder _;
struct List {
    template <typename... A> static List create(const A &...) { return List(); }
};
inline void warning(const char *) {}
[[noreturn]] inline void stop(const std::string &) { throw 0; }
inline void checkUserInterrupt() {}
}  // namespace Rcpp
