// Integration shim for builds without the fmt package: StormRuler only uses fmt::format to build
// log and exception strings (Storm/Crow/Base/Log.hpp:29-30, Exception.hpp:35-38); nothing on the
// Krylov path depends on the formatted text. Drop this directory from the include path when the
// real fmt is available.
#pragma once
#include <string>
#include <string_view>
namespace fmt {
template<class... Args>
inline std::string format(std::string_view message, const Args&...) {
  return std::string{message};
}
} // namespace fmt
