// Integration shim for builds without the spdlog package: StormRuler logs one line per solve
// (Storm/Solvers/Solver.hpp:144-145) and on aborts (Crow/Base/Assert.hpp:34-37). Drop this directory
// from the include path when the real spdlog is available.
#pragma once
#include <cstdio>
#include <string_view>
namespace spdlog {
template<class... A> inline void trace(std::string_view, const A&...) {}
template<class... A> inline void debug(std::string_view, const A&...) {}
template<class... A> inline void info(std::string_view, const A&...) {}
template<class... A> inline void warn(std::string_view, const A&...) {}
template<class... A> inline void error(std::string_view m, const A&...) {
  std::fprintf(stderr, "[storm error] %.*s\n", int(m.size()), m.data());
}
template<class... A> inline void critical(std::string_view m, const A&...) {
  std::fprintf(stderr, "[storm critical] %.*s\n", int(m.size()), m.data());
}
} // namespace spdlog
