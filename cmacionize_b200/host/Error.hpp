/*
 * Error.hpp — error convention of the host layer.
 *
 * The reference reports every error through cmac_error, which prints
 * "file:function():line: Error: message" to stderr and aborts
 * (/root/reference/src/Error.hpp:101-106).  The host layer raises the same
 * message as a C++ exception (cmi::Error); the command line program turns it
 * into the reference's print + abort, the C entry points of host_api.cpp turn it
 * into an error code (or abort when CMIB_ABORT_ON_ERROR=1, like libcmib).
 */
#pragma once
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace cmi {

class Error : public std::runtime_error {
public:
  explicit Error(const std::string &what) : std::runtime_error(what) {}
};

[[noreturn]] inline void raise_error(const char *file, const char *func, int line, const char *fmt, ...) {
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  char full[2400];
  snprintf(full, sizeof(full), "%s:%s():%d: Error:\n%s", file, func, line, msg);
  throw Error(full);
}

} // namespace cmi

#define cmi_error(...) ::cmi::raise_error(__FILE__, __func__, __LINE__, __VA_ARGS__)
