// TEST INFRASTRUCTURE (oracle build shim) -- not product code.
// Minimal stand-in for <fmt/format.h>: the reference only uses fmt::format to
// build log / exception strings (Storm/Crow/Base/Log.hpp:29-30,
// Exception.hpp:35-38). No arithmetic on the Krylov path lives in fmt, so the
// shim just returns the unformatted message.
#pragma once
#include <string>
#include <string_view>
namespace fmt {
template<class... Args>
inline std::string format(std::string_view message, const Args&...) {
  return std::string{message};
}
} // namespace fmt
