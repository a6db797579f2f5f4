// -*- C++ -*-
// Minimal stand-in for <RcppArmadillo.h>, TEST INFRASTRUCTURE ONLY (see oracle/arma_shim/armadillo).
// Just what rcpp-code/MultivarTV/src/{utils,solvers}.cpp use: Rcpp::Rcout, Rcpp::Nullable<T>, Rcpp::as<T>, Rcpp::List with
// Rcpp::Named entries, R_NilValue.  `// [[Rcpp::export]]` markers are comments and need nothing.
#pragma once
#include <any>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "armadillo"

struct R_NilValue_t {};
static const R_NilValue_t R_NilValue = R_NilValue_t();

namespace Rcpp {

static std::ostream &Rcout = std::cout;

template <typename T>
class Nullable {
 public:
  Nullable() : null_(true) {}
  Nullable(R_NilValue_t) : null_(true) {}
  Nullable(const T &v) : null_(false), value_(v) {}
  bool isNull() const { return null_; }
  bool isNotNull() const { return !null_; }
  const T &get() const { return value_; }

 private:
  bool null_;
  T value_;
};

template <typename T>
T as(const Nullable<T> &n) { return n.get(); }

struct NamedValue {
  std::string name;
  std::any value;
};
template <typename T>
NamedValue Named(const std::string &name, const T &value) { return NamedValue{name, std::any(value)}; }

class List {
 public:
  std::vector<NamedValue> items;
  List() {}
  template <typename... Args>
  static List create(const Args &...args) {
    List l;
    (l.items.push_back(args), ...);
    return l;
  }
  void push_back(const List &l) { items.push_back(NamedValue{"", std::any(l)}); }
  template <typename T>
  void push_back(const T &v) { items.push_back(NamedValue{"", std::any(v)}); }
  size_t size() const { return items.size(); }
  const std::any &operator[](const std::string &name) const {
    for (const auto &it : items)
      if (it.name == name) return it.value;
    throw std::out_of_range("Rcpp::List: no element named " + name);
  }
  const std::any &operator[](size_t i) const { return items.at(i).value; }
};

}  // namespace Rcpp
