// abi.cpp -- extern "C" surface: constructors, .cwipcdump IO, thin method wrappers, and the
// cwipc_cuda_* extensions.  Filters are in filters.cpp, logging in logging.cpp.
// ref: src/cwipc_util.cpp:412-870
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.hpp"
#include "pointcloud.hpp"

using namespace cwcu;

#ifndef CWIPC_CUDA_VERSION
#define CWIPC_CUDA_VERSION "cwipc_util_cuda-0.1-b200"
#endif

namespace {

// Every constructor checks the caller's API version first. ref: src/cwipc_util.cpp:663-670
bool version_ok(const char *who, uint64_t apiVersion, char **errorMessage) {
    if (apiVersion >= CWIPC_API_VERSION_OLD && apiVersion <= CWIPC_API_VERSION) return true;
    if (errorMessage) {
        char *msg = (char *)malloc(1024);
        snprintf(msg, 1024, "%s: incorrect apiVersion 0x%08" PRIx64 " expected 0x%08" PRIx64 "..0x%08" PRIx64 "", who, apiVersion, (uint64_t)CWIPC_API_VERSION_OLD,
                 (uint64_t)CWIPC_API_VERSION);
        *errorMessage = msg;
    }
    return false;
}

// Captures the first ERROR logged inside the scope into *errorMessage.
struct ErrorCapture {
    explicit ErrorCapture(char **errorMessage) { log_set_errorbuf(errorMessage); }
    ~ErrorCapture() { log_set_errorbuf(nullptr); }
};

cwipc_pointcloud *make_from_points(const char *who, const cwipc_point *points, size_t size, int npoint, uint64_t timestamp, bool sync) {
    if (npoint < 0 || (size_t)npoint * sizeof(cwipc_point) != size) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_util", "from_points: size and npoint inconsistent");
        log(CWIPC_LOG_LEVEL_ERROR, who, "cannot load points (size error?)");
        return nullptr;
    }
    return guarded<cwipc_pointcloud *>(who, nullptr, [&]() -> cwipc_pointcloud * {
        DevicePointcloud *pc = DevicePointcloud::from_host(points, (size_t)npoint, timestamp, sync);
        pc->set_exact_size(true); // the reference's cwipc_uncompressed_impl: copy_uncompressed(size != exact) is an error
        return pc;
    });
}

struct Timer {
    int dev;
    cudaStream_t stream;
    cudaEvent_t e0, e1;
};

} // namespace

extern "C" {

const char *cwipc_get_version(void) { return CWIPC_CUDA_VERSION; }

// ---- constructors ----------------------------------------------------------------------------
cwipc_pointcloud *cwipc_from_points(struct cwipc_point *points, size_t size, int npoint, uint64_t timestamp, char **errorMessage, uint64_t apiVersion) {
    if (!version_ok("cwipc_from_points", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    return make_from_points("cwipc_from_points", points, size, npoint, timestamp, true);
}

cwipc_pointcloud *cwipc_cuda_from_points_async(struct cwipc_point *points, size_t size, int npoint, uint64_t timestamp, char **errorMessage, uint64_t apiVersion) {
    if (!version_ok("cwipc_cuda_from_points_async", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    return make_from_points("cwipc_cuda_from_points_async", points, size, npoint, timestamp, false);
}

cwipc_pointcloud *cwipc_from_packet(uint8_t *packet, size_t size, char **errorMessage, uint64_t apiVersion) {
    if (!version_ok("cwipc_from_packet", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    cwipc_cwipcdump_header hdr;
    if (packet == nullptr || size < sizeof(hdr)) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_util", "cwipc_from_packet: packet too small");
        return nullptr;
    }
    memcpy(&hdr, packet, sizeof(hdr));
    if (memcmp(hdr.hdr, CWIPC_CWIPCDUMP_HEADER, 4) != 0 || hdr.magic != CWIPC_CWIPCDUMP_VERSION) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_util", "cwipc_from_packet: incorrect packet header or version");
        return nullptr;
    }
    const size_t dataSize = size - sizeof(hdr);
    const size_t npoint = hdr.size / sizeof(cwipc_point);
    if (npoint * sizeof(cwipc_point) != dataSize) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_util", "cwipc_from_packet: inconsistent dataSize");
        return nullptr;
    }
    cwipc_pointcloud *rv = make_from_points("cwipc_from_packet", reinterpret_cast<const cwipc_point *>(packet + sizeof(hdr)), dataSize, (int)npoint, hdr.timestamp, true);
    if (rv) rv->_set_cellsize(hdr.cellsize);
    return rv;
}

// ---- .cwipcdump files: 32-byte header + raw points.  ref: src/cwipc_util.cpp:499-641 ---------
cwipc_pointcloud *cwipc_read_debugdump(const char *filename, char **errorMessage, uint64_t apiVersion) {
    if (!version_ok("cwipc_read_debugdump", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    FILE *fp = fopen(filename, "rb");
    if (!fp) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", std::string("Cannot open file: ") + filename);
        return nullptr;
    }
    cwipc_pointcloud *rv = nullptr;
    cwipc_cwipcdump_header hdr;
    std::vector<cwipc_point> data;
    if (fread(&hdr, 1, sizeof(hdr), fp) != sizeof(hdr)) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", std::string("Cannot read pointcloud dumpfile header: ") + filename);
    } else if (memcmp(hdr.hdr, CWIPC_CWIPCDUMP_HEADER, 4) != 0) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", std::string("Pointcloud dumpfile header incorrect: ") + filename);
    } else if (hdr.magic != CWIPC_CWIPCDUMP_VERSION) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", std::string("Pointcloud dumpfile version incorrect: ") + filename);
    } else if ((hdr.size / sizeof(cwipc_point)) * sizeof(cwipc_point) != hdr.size) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", "Pointcloud dumpfile datasize inconsistent");
    } else {
        const size_t npoint = hdr.size / sizeof(cwipc_point);
        data.resize(npoint);
        if (npoint && fread(data.data(), 1, hdr.size, fp) != hdr.size) {
            log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read_debugdump", "Could not read point data of correct size");
        } else {
            rv = make_from_points("cwipc_read_debugdump", data.data(), hdr.size, (int)npoint, hdr.timestamp, true);
            if (rv) rv->_set_cellsize(hdr.cellsize);
        }
    }
    fclose(fp);
    return rv;
}

int cwipc_write_debugdump(const char *filename, cwipc_pointcloud *pc, char **errorMessage) {
    ErrorCapture cap(errorMessage);
    if (pc == nullptr) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_write_debugdump", "NULL pointcloud");
        return -1;
    }
    const size_t dataSize = pc->get_uncompressed_size();
    std::vector<cwipc_point> data(dataSize / sizeof(cwipc_point));
    const int npoint = dataSize ? pc->copy_uncompressed(data.data(), dataSize) : 0;
    if (npoint < 0) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_write_debugdump", "Cannot copy points, size=" + std::to_string(dataSize));
        return -1;
    }
    FILE *fp = fopen(filename, "wb");
    if (!fp) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_write_debugdump", std::string("Cannot open output file: ") + filename);
        return -1;
    }
    cwipc_cwipcdump_header hdr;
    memset(&hdr, 0, sizeof(hdr));
    memcpy(hdr.hdr, CWIPC_CWIPCDUMP_HEADER, 4);
    hdr.magic = CWIPC_CWIPCDUMP_VERSION;
    hdr.timestamp = pc->timestamp();
    hdr.cellsize = pc->cellsize();
    hdr.size = dataSize;
    int status = 0;
    if (fwrite(&hdr, sizeof(hdr), 1, fp) != 1 || fwrite(data.data(), sizeof(cwipc_point), (size_t)npoint, fp) != (size_t)npoint) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_write_debugdump", "Cannot write point data, nPoint=" + std::to_string(npoint));
        status = -1;
    }
    fclose(fp);
    return status;
}

// ---- method wrappers.  ref: src/cwipc_util.cpp:731-797 ----------------------------------------
void cwipc_pointcloud_free(cwipc_pointcloud *pc) { pc->free(); }
cwipc_pointcloud *cwipc_pointcloud__shallowcopy(cwipc_pointcloud *pc) { return pc->_shallowcopy(); }
uint64_t cwipc_pointcloud_timestamp(cwipc_pointcloud *pc) { return pc->timestamp(); }
float cwipc_pointcloud_cellsize(cwipc_pointcloud *pc) { return pc->cellsize(); }
void cwipc_pointcloud__set_cellsize(cwipc_pointcloud *pc, float cellsize) { pc->_set_cellsize(cellsize); }
void cwipc_pointcloud__set_timestamp(cwipc_pointcloud *pc, uint64_t timestamp) { pc->_set_timestamp(timestamp); }
int cwipc_pointcloud_count(cwipc_pointcloud *pc) { return pc->count(); }
size_t cwipc_pointcloud_get_uncompressed_size(cwipc_pointcloud *pc) { return pc->get_uncompressed_size(); }
int cwipc_pointcloud_copy_uncompressed(cwipc_pointcloud *pc, struct cwipc_point *pointbuf, size_t size) { return pc->copy_uncompressed(pointbuf, size); }
size_t cwipc_pointcloud_copy_packet(cwipc_pointcloud *pc, uint8_t *packet, size_t size) { return pc->copy_packet(packet, size); }
cwipc_metadata *cwipc_pointcloud_access_metadata(cwipc_pointcloud *pc) { return pc->access_metadata(); }

void cwipc_metadata__move(cwipc_metadata *src, cwipc_metadata *dest) { src->_move(dest); }
int cwipc_metadata_count(cwipc_metadata *collection) { return collection->count(); }
const char *cwipc_metadata_name(cwipc_metadata *collection, int idx) { return collection->name(idx).c_str(); }
const char *cwipc_metadata_description(cwipc_metadata *collection, int idx) { return collection->description(idx).c_str(); }
void *cwipc_metadata_pointer(cwipc_metadata *collection, int idx) { return collection->pointer(idx); }
size_t cwipc_metadata_size(cwipc_metadata *collection, int idx) { return collection->size(idx); }

// ref: src/cwipc_util.cpp:799-870
bool cwipc_activesource_start(cwipc_activesource *src) { return src->start(); }
void cwipc_activesource_stop(cwipc_activesource *src) { src->stop(); }
cwipc_pointcloud *cwipc_source_get(cwipc_source *src) { return src->get(); }
void cwipc_source_free(cwipc_source *src) { src->free(); }
bool cwipc_source_eof(cwipc_source *src) { return src->eof(); }
bool cwipc_source_available(cwipc_source *src, bool wait) { return src->available(wait); }
void cwipc_activesource_request_metadata(cwipc_activesource *src, const char *name) { src->request_metadata(name); }
bool cwipc_activesource_is_metadata_requested(cwipc_activesource *src, const char *name) { return src->is_metadata_requested(name); }
bool cwipc_activesource_reload_config(cwipc_activesource *src, const char *configFile) { return src->reload_config(configFile); }
size_t cwipc_activesource_get_config(cwipc_activesource *src, char *buffer, size_t size) { return src->get_config(buffer, size); }
bool cwipc_activesource_seek(cwipc_activesource *src, uint64_t timestamp) { return src->seek(timestamp); }
int cwipc_activesource_maxtile(cwipc_activesource *src) { return src->maxtile(); }
bool cwipc_activesource_get_tileinfo(cwipc_activesource *src, int tilenum, struct cwipc_tileinfo *tileinfo) { return src->get_tileinfo(tilenum, tileinfo); }
bool cwipc_activesource_auxiliary_operation(cwipc_activesource *src, const char *op, const void *inbuf, size_t insize, void *outbuf, size_t outsize) {
    return src->auxiliary_operation(std::string(op), inbuf, insize, outbuf, outsize);
}
void cwipc_sink_free(cwipc_sink *sink) { sink->free(); }
bool cwipc_sink_feed(cwipc_sink *sink, cwipc_pointcloud *pc, bool clear) { return sink->feed(pc, clear); }
bool cwipc_sink_caption(cwipc_sink *sink, const char *caption) { return sink->caption(caption); }
char cwipc_sink_interact(cwipc_sink *sink, const char *prompt, const char *responses, int32_t millis) { return sink->interact(prompt, responses, millis); }

// ---- out-of-scope factories: same failure shape as the reference's non-GUI / unknown-camera builds
cwipc_activesource *cwipc_capturer(const char *configFilename, char **errorMessage, uint64_t apiVersion) {
    (void)configFilename;
    if (!version_ok("cwipc_capturer", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    log(CWIPC_LOG_LEVEL_ERROR, "cwipc_capturer", "no capturers are registered in libcwipc_util_cuda (camera plumbing is out of scope)");
    return nullptr;
}

cwipc_sink *cwipc_window(const char *title, char **errorMessage, uint64_t apiVersion) {
    (void)title;
    if (!version_ok("cwipc_window", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    log(CWIPC_LOG_LEVEL_ERROR, "cwipc_window", "libcwipc_util_cuda is built without GUI support");
    return nullptr;
}

cwipc_activesource *cwipc_proxy(const char *host, int port, char **errorMessage, uint64_t apiVersion) {
    (void)host;
    (void)port;
    if (!version_ok("cwipc_proxy", apiVersion, errorMessage)) return nullptr;
    ErrorCapture cap(errorMessage);
    log(CWIPC_LOG_LEVEL_ERROR, "cwipc_proxy", "the TCP proxy source is not part of libcwipc_util_cuda");
    return nullptr;
}

// ---- cwipc_cuda_* extensions -------------------------------------------------------------------
int cwipc_cuda_device_count(void) { return device_count(); }
int cwipc_cuda_set_device(int device) { return set_current_device(device) ? 0 : -1; }
int cwipc_cuda_get_device(void) { return current_device(); }

int cwipc_cuda_synchronize(void) {
    return guarded<int>("cwipc_cuda_synchronize", -1, [&] {
        const int dev = current_device();
        DeviceGuard g(dev);
        stream_sync(thread_stream(dev));
        return 0;
    });
}

void *cwipc_cuda_host_alloc(size_t size) {
    return guarded<void *>("cwipc_cuda_host_alloc", nullptr, [&] {
        if (device_count() <= 0) throw CudaError{cudaErrorNoDevice, "no CUDA device"};
        DeviceGuard g(current_device());
        void *p = nullptr;
        CWCU_CHECK(cudaHostAlloc(&p, size ? size : 16, cudaHostAllocPortable));
        return p;
    });
}

void cwipc_cuda_host_free(void *ptr) {
    if (ptr) (void)cudaFreeHost(ptr);
}

int cwipc_cuda_pointcloud_device(cwipc_pointcloud *pc) {
    auto *mine = dynamic_cast<DevicePointcloud *>(pc);
    return (mine && mine->storage()) ? mine->storage()->dev : -1;
}

const void *cwipc_cuda_pointcloud_device_ptr(cwipc_pointcloud *pc) {
    auto *mine = dynamic_cast<DevicePointcloud *>(pc);
    if (!mine || !mine->storage()) return nullptr;
    // make the contents valid for any consumer: wait for the producing work
    (void)cudaEventSynchronize(mine->storage()->ready);
    return mine->storage()->d_pts;
}

void *cwipc_cuda_timer_create(void) {
    return guarded<void *>("cwipc_cuda_timer_create", nullptr, [&]() -> void * {
        const int dev = current_device();
        DeviceGuard g(dev);
        Timer *t = new Timer();
        t->dev = dev;
        t->stream = thread_stream(dev);
        CWCU_CHECK(cudaEventCreate(&t->e0));
        CWCU_CHECK(cudaEventCreate(&t->e1));
        return t;
    });
}

void cwipc_cuda_timer_destroy(void *timer) {
    Timer *t = static_cast<Timer *>(timer);
    if (!t) return;
    (void)cudaEventDestroy(t->e0);
    (void)cudaEventDestroy(t->e1);
    delete t;
}

void cwipc_cuda_timer_start(void *timer) {
    Timer *t = static_cast<Timer *>(timer);
    DeviceGuard g(t->dev);
    t->stream = thread_stream(t->dev);
    (void)cudaEventRecord(t->e0, t->stream);
}

void cwipc_cuda_timer_stop(void *timer) {
    Timer *t = static_cast<Timer *>(timer);
    DeviceGuard g(t->dev);
    (void)cudaEventRecord(t->e1, thread_stream(t->dev));
}

float cwipc_cuda_timer_elapsed_ms(void *timer) { return cwipc_cuda_timer_span_ms(timer, timer); }

float cwipc_cuda_timer_span_ms(void *timer_a, void *timer_b) {
    Timer *a = static_cast<Timer *>(timer_a), *b = static_cast<Timer *>(timer_b);
    DeviceGuard g(a->dev);
    float ms = -1.f;
    if (cudaEventSynchronize(b->e1) != cudaSuccess || cudaEventSynchronize(a->e0) != cudaSuccess || cudaEventElapsedTime(&ms, a->e0, b->e1) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1.f;
    }
    return ms;
}

uint64_t cwipc_cuda_kernel_launches(void) { return g_kernel_launches.load(); }
void cwipc_cuda_profile_enable(int on) { profile_enable(on != 0); }
void cwipc_cuda_profile_reset(void) { profile_reset(); }
size_t cwipc_cuda_profile_report(char *buf, size_t size) {
    const std::string js = profile_report_json();
    if (buf && size > 0) {
        const size_t n = js.size() < size - 1 ? js.size() : size - 1;
        memcpy(buf, js.data(), n);
        buf[n] = 0;
    }
    return js.size() + 1;
}

void cwipc_cuda_flush_l2(void) {
    guarded<int>("cwipc_cuda_flush_l2", 0, [&] {
        const int dev = current_device();
        DeviceGuard g(dev);
        flush_l2(dev, thread_stream(dev));
        return 0;
    });
}

int cwipc_cuda_trim(void) {
    return guarded<int>("cwipc_cuda_trim", -1, [&] {
        trim_device(current_device());
        return 0;
    });
}

} // extern "C"
