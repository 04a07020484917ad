// Tiny expression evaluator for boundary-condition strings in x, y, z, t -- the role Pamgen RTC
// plays in the reference (src/control.cpp:15-23,104-120).  Grammar: + - * / ^, unary -, (),
// numbers, the variables x y z t, pi, and sin cos tan atan exp sqrt abs log pow(a,b).
#pragma once
#include <cctype>
#include <cmath>
#include <stdexcept>
#include <string>

namespace c8host {

class Expr {
 public:
  explicit Expr(std::string s) : src_(std::move(s)) {}
  double operator()(double x, double y, double z, double t) const {
    P p{src_.c_str(), x, y, z, t};
    const double v = p.sum();
    p.ws();
    if (*p.s) throw std::runtime_error("expression: trailing characters in '" + src_ + "'");
    return v;
  }
  const std::string& str() const { return src_; }

 private:
  std::string src_;
  struct P {
    const char* s;
    double x, y, z, t;
    void ws() { while (*s && std::isspace((unsigned char)*s)) ++s; }
    bool eat(char c) { ws(); if (*s == c) { ++s; return true; } return false; }
    double sum() {
      double v = prod();
      for (;;) {
        if (eat('+')) v += prod();
        else if (eat('-')) v -= prod();
        else return v;
      }
    }
    double prod() {
      double v = unary();
      for (;;) {
        if (eat('*')) v *= unary();
        else if (eat('/')) v /= unary();
        else return v;
      }
    }
    double unary() {
      if (eat('-')) return -unary();
      if (eat('+')) return unary();
      return power();
    }
    double power() {
      double b = atom();
      if (eat('^')) return std::pow(b, unary());
      return b;
    }
    double atom() {
      ws();
      if (eat('(')) { double v = sum(); if (!eat(')')) throw std::runtime_error("expression: missing )"); return v; }
      if (std::isdigit((unsigned char)*s) || *s == '.') {
        char* end = nullptr;
        double v = std::strtod(s, &end);
        s = end;
        return v;
      }
      if (std::isalpha((unsigned char)*s)) {
        std::string id;
        while (std::isalnum((unsigned char)*s) || *s == '_') id += *s++;
        if (id == "x") return x;
        if (id == "y") return y;
        if (id == "z") return z;
        if (id == "t") return t;
        if (id == "pi") return 3.14159265358979323846;
        if (!eat('(')) throw std::runtime_error("expression: unknown identifier " + id);
        double a = sum(), b = 0.0;
        if (id == "pow") { if (!eat(',')) throw std::runtime_error("pow needs 2 args"); b = sum(); }
        if (!eat(')')) throw std::runtime_error("expression: missing )");
        if (id == "sin") return std::sin(a);
        if (id == "cos") return std::cos(a);
        if (id == "tan") return std::tan(a);
        if (id == "atan") return std::atan(a);
        if (id == "asin") return std::asin(a);
        if (id == "acos") return std::acos(a);
        if (id == "tanh") return std::tanh(a);
        if (id == "exp") return std::exp(a);
        if (id == "sqrt") return std::sqrt(a);
        if (id == "abs" || id == "fabs") return std::fabs(a);
        if (id == "log") return std::log(a);
        if (id == "pow") return std::pow(a, b);
        throw std::runtime_error("expression: unknown function " + id);
      }
      throw std::runtime_error(std::string("expression: unexpected '") + *s + "'");
    }
  };
};

}  // namespace c8host
