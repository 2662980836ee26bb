// logging.cpp -- the reference's logging contract, restated: level filter, stderr / file /
// callback sinks, CWIPC_LOGGING=LEVEL[:file], and capture of the first ERROR of a call into the
// caller's `char **errorMessage`.
// ref: src/logging.cpp:18-143, include/cwipc_util/internal/logging.hpp:7-22
//
// Differences from the reference: state is guarded by a mutex and the error buffer is per thread,
// because this library is entered concurrently from several host threads (one stream each).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <mutex>

#include "runtime.hpp"

namespace cwcu {

namespace {

struct LogState {
    std::mutex mu;
    bool configured = false;
    cwipc_log_level level = CWIPC_LOG_LEVEL_WARNING;
    cwipc_log_callback_t callback = nullptr;
    std::ostream *stream = nullptr; // set when CWIPC_LOGGING is present
    bool to_stderr = true, to_file = false, to_callback = false;
    time_t t0 = 0;
};

LogState &state() {
    static LogState *s = new LogState(); // leaked on purpose: logging may run during exit
    return *s;
}

thread_local char **t_errorbuf = nullptr;

cwipc_log_level parse_level(const std::string &name) {
    static const struct { const char *n; cwipc_log_level l; } table[] = {
        {"NONE", CWIPC_LOG_LEVEL_NONE}, {"ERROR", CWIPC_LOG_LEVEL_ERROR}, {"WARNING", CWIPC_LOG_LEVEL_WARNING},
        {"TRACE", CWIPC_LOG_LEVEL_TRACE}, {"DEBUG", CWIPC_LOG_LEVEL_DEBUG}};
    for (auto &e : table)
        if (name == e.n) return e.l;
    return CWIPC_LOG_LEVEL_WARNING;
}

const char *level_name(cwipc_log_level level) {
    switch (level) {
    case CWIPC_LOG_LEVEL_ERROR: return "Error";
    case CWIPC_LOG_LEVEL_WARNING: return "Warning";
    case CWIPC_LOG_LEVEL_TRACE: return "Trace";
    case CWIPC_LOG_LEVEL_DEBUG: return "Debug";
    default: return "Unknown-level";
    }
}

// caller holds st.mu
void configure_from_env(LogState &st) {
    if (st.configured) return;
    st.configured = true;
    const char *env = getenv("CWIPC_LOGGING");
    if (!env) return;
    std::string spec(env), file;
    const size_t colon = spec.find(':');
    if (colon != std::string::npos) {
        file = spec.substr(colon + 1);
        spec.resize(colon);
    }
    st.level = parse_level(spec);
    st.stream = file.empty() ? &std::cerr : new std::ofstream(file, std::ios::out | std::ios::app);
    st.to_stderr = false;
    st.to_file = true;
}

} // namespace

void log(cwipc_log_level level, const std::string &module, const std::string &message) {
    LogState &st = state();
    cwipc_log_callback_t cb = nullptr;
    std::string line;
    {
        std::lock_guard<std::mutex> lk(st.mu);
        configure_from_env(st);
        if (level > st.level) return;
        line = module + ": " + level_name(level) + ": " + message;
        if (st.t0 == 0) st.t0 = time(nullptr);
        const std::string stamp = "t=" + std::to_string((long long)(time(nullptr) - st.t0)) + ": ";
        // first error of the call wins; like the reference this string is handed to the caller
        if (t_errorbuf && level == CWIPC_LOG_LEVEL_ERROR && *t_errorbuf == nullptr) *t_errorbuf = strdup(line.c_str());
        if (st.to_stderr) std::cerr << stamp << line << std::endl;
        if (st.to_file && st.stream) {
            (*st.stream) << stamp << line << std::endl;
            st.stream->flush();
        }
        if (st.to_callback) cb = st.callback;
    }
    if (cb) cb((int)level, line.c_str()); // outside the lock: the callback may log again
}

void log_set_errorbuf(char **errorbuf) { t_errorbuf = errorbuf; }

cwipc_log_level log_get_level() {
    LogState &st = state();
    std::lock_guard<std::mutex> lk(st.mu);
    configure_from_env(st);
    return st.level;
}

void log_configure(int level, cwipc_log_callback_t callback) {
    LogState &st = state();
    bool debug;
    {
        std::lock_guard<std::mutex> lk(st.mu);
        configure_from_env(st);
        if (level != CWIPC_LOG_LEVEL_NONE) st.level = (cwipc_log_level)level;
        st.callback = callback;
        st.to_callback = callback != nullptr;
        st.to_stderr = callback ? false : !st.to_file;
        debug = st.level >= CWIPC_LOG_LEVEL_DEBUG;
    }
    if (debug) log(CWIPC_LOG_LEVEL_DEBUG, "logging", "Logging configured, (int)callback=" + std::to_string((intptr_t)callback));
}

} // namespace cwcu

namespace cwcu { void log_configure(int level, cwipc_log_callback_t callback); }

extern "C" {

void cwipc_log_configure(int level, cwipc_log_callback_t callback) { cwcu::log_configure(level, callback); }

void _cwipc_log_emit(int level, const char *module, const char *message) {
    cwcu::log((cwipc_log_level)level, module ? module : "", message ? message : "");
}

}
