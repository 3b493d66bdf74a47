// FynException + THROW_EXCEPTION_ARGS: error convention of the host engine.
// Mirrors the reference's convention (fyusenet/common/fynexception.h:24-25,75-107): exceptions carry
// a printf-formatted message plus the throwing function / file / line.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <exception>
#include <string>

namespace fyusion {

class FynException : public std::exception {
 public:
    FynException() = default;
    FynException(const char *function, const char *file, int line, const char *fmt, ...) {
        char buf[2048];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        message_ = std::string(buf) + " [" + (function ? function : "?") + " @ " + (file ? file : "?") + ":" +
                   std::to_string(line) + "]";
    }
    const char *what() const noexcept override { return message_.c_str(); }

 protected:
    std::string message_;
};

}  // namespace fyusion

#define THROW_EXCEPTION_ARGS(cls, ...) throw cls(__FUNCTION__, __FILE__, __LINE__, __VA_ARGS__)
