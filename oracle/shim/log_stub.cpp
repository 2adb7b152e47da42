// Definitions for the reference's Log class (Log.h is used as-is; the real Log.cpp pulls in SDL and the
// asset loader). TEST INFRASTRUCTURE ONLY.
#include "Log.h"
#include <cstdio>
namespace Atlas {
    std::vector<Log::Entry> Log::entries;
    std::mutex Log::mutex;
    void Log::Message(const std::string&, int32_t) {}
    void Log::Warning(const std::string& m, int32_t) { if (getenv("ATLAS_REF_VERBOSE")) fprintf(stderr, "[ref warning] %s\n", m.c_str()); }
    void Log::Error(const std::string& m, int32_t) { fprintf(stderr, "[ref error] %s\n", m.c_str()); }
}
