// Oracle shim: glog is absent.  LOG(severity) << ... counts ERROR messages (so the C API of the _ref build can
// report "the reference logged an error and returned early", src/ORBextractor.cpp:1085-1088,1183-1186) and
// otherwise swallows the text.
#ifndef SLAMB200_ORACLE_SHIM_GLOG
#define SLAMB200_ORACLE_SHIM_GLOG
#include <sstream>
namespace shim_glog {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
extern int error_count;
struct Sink {
    explicit Sink(int sev) { if (sev >= ERROR) ++error_count; }
    template <typename T> Sink &operator<<(const T &) { return *this; }
    Sink &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
}
#define LOG(sev) ::shim_glog::Sink(::shim_glog::sev)
#endif
