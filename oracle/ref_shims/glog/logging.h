// Minimal stand-in for <glog/logging.h>, used ONLY to compile the unmodified
// reference sources under /root/reference into oracle/_ref/ (test
// infrastructure; see oracle/Makefile).  The reference vendors glog as a
// cmake sub-project whose logging.h is generated; this header supplies the
// macro surface the reference's renderer sources use (LOG/VLOG/CHECK*/DCHECK*)
// with glog's observable semantics: INFO/WARNING are dropped unless
// FLAGS_minloglevel / FLAGS_logtostderr ask for them, ERROR goes to stderr,
// FATAL and failed CHECKs abort.
#ifndef IILE_ORACLE_GLOG_SHIM_H
#define IILE_ORACLE_GLOG_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char *) {}
}  // namespace google

extern int FLAGS_stderrthreshold;
extern int FLAGS_minloglevel;
extern int FLAGS_v;
extern bool FLAGS_logtostderr;
extern std::string FLAGS_log_dir;

namespace glogshim {

class Message {
  public:
    Message(const char *file, int line, int severity, bool live)
        : severity_(severity), live_(live) {
        if (live_) ss_ << "[" << "IWEF"[severity] << " " << file << ":" << line << "] ";
    }
    ~Message() {
        if (live_) {
            ss_ << "\n";
            std::string s = ss_.str();
            fwrite(s.data(), 1, s.size(), stderr);
        }
        if (severity_ == google::GLOG_FATAL) {
            fflush(stderr);
            abort();
        }
    }
    std::ostream &stream() { return ss_; }

  private:
    std::ostringstream ss_;
    int severity_;
    bool live_;
};

struct Voidify {
    void operator&(std::ostream &) {}
};

inline bool Live(int severity) {
    if (severity >= google::GLOG_FATAL) return true;
    if (severity < FLAGS_minloglevel) return false;
    return FLAGS_logtostderr || severity >= FLAGS_stderrthreshold;
}

template <typename T>
T CheckNotNull(const char *file, int line, const char *expr, T &&t) {
    if (t == nullptr) {
        Message(file, line, google::GLOG_FATAL, true).stream()
            << "Check failed: '" << expr << "' Must be non NULL";
    }
    return std::forward<T>(t);
}

}  // namespace glogshim

#define GLOGSHIM_SEV_INFO google::GLOG_INFO
#define GLOGSHIM_SEV_WARNING google::GLOG_WARNING
#define GLOGSHIM_SEV_ERROR google::GLOG_ERROR
#define GLOGSHIM_SEV_FATAL google::GLOG_FATAL

#define LOG(sev)                                                      \
    !glogshim::Live(GLOGSHIM_SEV_##sev)                               \
        ? (void)0                                                     \
        : glogshim::Voidify() &                                       \
              glogshim::Message(__FILE__, __LINE__, GLOGSHIM_SEV_##sev, true).stream()

#define VLOG(n)                                                       \
    !((n) <= FLAGS_v && glogshim::Live(google::GLOG_INFO))            \
        ? (void)0                                                     \
        : glogshim::Voidify() &                                       \
              glogshim::Message(__FILE__, __LINE__, google::GLOG_INFO, true).stream()

#define CHECK(cond)                                                   \
    (cond) ? (void)0                                                  \
           : glogshim::Voidify() &                                    \
                 glogshim::Message(__FILE__, __LINE__, google::GLOG_FATAL, true).stream() \
                     << "Check failed: " #cond " "

#define GLOGSHIM_CHECK_OP(a, op, b)                                   \
    ((a)op(b)) ? (void)0                                              \
               : glogshim::Voidify() &                                \
                     glogshim::Message(__FILE__, __LINE__, google::GLOG_FATAL, true).stream() \
                         << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "

#define CHECK_EQ(a, b) GLOGSHIM_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) GLOGSHIM_CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) GLOGSHIM_CHECK_OP(a, <, b)
#define CHECK_LE(a, b) GLOGSHIM_CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) GLOGSHIM_CHECK_OP(a, >, b)
#define CHECK_GE(a, b) GLOGSHIM_CHECK_OP(a, >=, b)
#define CHECK_NEAR(a, b, eps) CHECK(((a) - (b)) <= (eps) && ((b) - (a)) <= (eps))
#define CHECK_NOTNULL(p) glogshim::CheckNotNull(__FILE__, __LINE__, #p, (p))

// NDEBUG build (the reference's Release configuration): DCHECKs compile to
// dead code that still type-checks its stream operands, as glog's do.
#define GLOGSHIM_DEAD(cond) \
    true ? (void)0 : glogshim::Voidify() & glogshim::Message(__FILE__, __LINE__, 0, false).stream()
#ifdef NDEBUG
#define DCHECK(cond) GLOGSHIM_DEAD(cond)
#define DCHECK_EQ(a, b) GLOGSHIM_DEAD((a) == (b))
#define DCHECK_NE(a, b) GLOGSHIM_DEAD((a) != (b))
#define DCHECK_LT(a, b) GLOGSHIM_DEAD((a) < (b))
#define DCHECK_LE(a, b) GLOGSHIM_DEAD((a) <= (b))
#define DCHECK_GT(a, b) GLOGSHIM_DEAD((a) > (b))
#define DCHECK_GE(a, b) GLOGSHIM_DEAD((a) >= (b))
#else
#define DCHECK(cond) CHECK(cond)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#endif

#endif  // IILE_ORACLE_GLOG_SHIM_H
