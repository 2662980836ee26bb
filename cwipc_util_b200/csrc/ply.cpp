// ply.cpp -- minimal PLY reader/writer behind cwipc_read / cwipc_write / cwipc_write_ext so that the
// cwipc_downsample / cwipc_remove_outliers / cwipc_tilefilter apps are PLY-in / PLY-out drop-ins.
// The reference delegates to pcl::PLYReader / pcl::PLYWriter (src/cwipc_util.cpp:432-497); the file
// layout restated here is what PCL produces for a cloud with fields x,y,z,rgba: vertex properties
// float x,y,z + uchar red,green,blue,alpha (the tile number travels in alpha), ASCII by default and
// binary_little_endian with CWIPC_FLAG_BINARY.  Host-side IO only: no kernels involved.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "pointcloud.hpp"

using namespace cwcu;

namespace {

enum PlyType { T_INT8, T_UINT8, T_INT16, T_UINT16, T_INT32, T_UINT32, T_FLOAT32, T_FLOAT64, T_BAD };

PlyType parse_type(const std::string &t) {
    if (t == "char" || t == "int8") return T_INT8;
    if (t == "uchar" || t == "uint8") return T_UINT8;
    if (t == "short" || t == "int16") return T_INT16;
    if (t == "ushort" || t == "uint16") return T_UINT16;
    if (t == "int" || t == "int32") return T_INT32;
    if (t == "uint" || t == "uint32") return T_UINT32;
    if (t == "float" || t == "float32") return T_FLOAT32;
    if (t == "double" || t == "float64") return T_FLOAT64;
    return T_BAD;
}

size_t type_size(PlyType t) {
    static const size_t sz[] = {1, 1, 2, 2, 4, 4, 4, 8, 0};
    return sz[t];
}

double read_binary(const uint8_t *p, PlyType t) {
    switch (t) {
    case T_INT8: return *(const int8_t *)p;
    case T_UINT8: return *p;
    case T_INT16: { int16_t v; memcpy(&v, p, 2); return v; }
    case T_UINT16: { uint16_t v; memcpy(&v, p, 2); return v; }
    case T_INT32: { int32_t v; memcpy(&v, p, 4); return v; }
    case T_UINT32: { uint32_t v; memcpy(&v, p, 4); return v; }
    case T_FLOAT32: { float v; memcpy(&v, p, 4); return v; }
    case T_FLOAT64: { double v; memcpy(&v, p, 8); return v; }
    default: return 0;
    }
}

struct Prop {
    std::string name;
    PlyType type;
    size_t offset;
};

// which point field a property feeds: 0..2 xyz, 3..5 rgb, 6 alpha/tile, 7 packed rgb(a), -1 ignored
int field_of(const std::string &n) {
    if (n == "x") return 0;
    if (n == "y") return 1;
    if (n == "z") return 2;
    if (n == "red" || n == "r") return 3;
    if (n == "green" || n == "g") return 4;
    if (n == "blue" || n == "b") return 5;
    if (n == "alpha" || n == "a" || n == "tile") return 6;
    if (n == "rgba" || n == "rgb") return 7;
    return -1;
}

void assign_packed(cwipc_point &pt, uint32_t rgba) { // a<<24 | r<<16 | g<<8 | b  (include/cwipc_util/api_pcl.h:20-55)
    pt.tile = (uint8_t)(rgba >> 24);
    pt.r = (uint8_t)(rgba >> 16);
    pt.g = (uint8_t)(rgba >> 8);
    pt.b = (uint8_t)rgba;
}

void assign(cwipc_point &pt, int field, double v) {
    switch (field) {
    case 0: pt.x = (float)v; break;
    case 1: pt.y = (float)v; break;
    case 2: pt.z = (float)v; break;
    case 3: pt.r = (uint8_t)v; break;
    case 4: pt.g = (uint8_t)v; break;
    case 5: pt.b = (uint8_t)v; break;
    case 6: pt.tile = (uint8_t)v; break;
    case 7: assign_packed(pt, (uint32_t)v); break;
    default: break;
    }
}

bool read_ply(const char *filename, std::vector<cwipc_point> &points, std::string &err) {
    FILE *fp = fopen(filename, "rb");
    if (!fp) {
        err = std::string("cannot open: ") + strerror(errno);
        return false;
    }
    char line[1024];
    bool binary = false, in_vertex = false, seen_vertex = false, header_done = false;
    size_t nvertex = 0, skip_before = 0, stride = 0;
    std::vector<Prop> props;
    bool first = true, other_before = false;
    while (fgets(line, sizeof(line), fp)) {
        std::istringstream ls(line);
        std::string tok;
        ls >> tok;
        if (first) {
            first = false;
            if (tok != "ply") { err = "not a PLY file"; fclose(fp); return false; }
            continue;
        }
        if (tok == "format") {
            std::string f;
            ls >> f;
            if (f == "ascii") binary = false;
            else if (f == "binary_little_endian") binary = true;
            else { err = "unsupported PLY format " + f; fclose(fp); return false; }
        } else if (tok == "element") {
            std::string name;
            size_t cnt = 0;
            ls >> name >> cnt;
            in_vertex = (name == "vertex");
            if (in_vertex) { seen_vertex = true; nvertex = cnt; }
            else if (!seen_vertex) { other_before = true; skip_before += cnt; }
        } else if (tok == "property") {
            std::string t, name;
            ls >> t;
            if (in_vertex) {
                if (t == "list") { err = "list property in vertex element"; fclose(fp); return false; }
                ls >> name;
                PlyType pt = parse_type(t);
                if (pt == T_BAD) { err = "unknown property type " + t; fclose(fp); return false; }
                props.push_back(Prop{name, pt, stride});
                stride += type_size(pt);
            }
        } else if (tok == "end_header") {
            header_done = true;
            break;
        }
    }
    if (!header_done || !seen_vertex) { err = "PLY header incomplete"; fclose(fp); return false; }
    if (other_before && binary) { err = "binary PLY with elements before vertex is not supported"; fclose(fp); return false; }
    points.assign(nvertex, cwipc_point{0, 0, 0, 0, 0, 0, 0});
    std::vector<int> fields(props.size());
    for (size_t i = 0; i < props.size(); i++) fields[i] = field_of(props[i].name);
    bool ok = true;
    if (binary) {
        std::vector<uint8_t> row(stride ? stride : 1);
        for (size_t v = 0; v < nvertex && ok; v++) {
            if (fread(row.data(), 1, stride, fp) != stride) { ok = false; break; }
            for (size_t i = 0; i < props.size(); i++) {
                if (fields[i] == 7 && props[i].type == T_FLOAT32) { // legacy "float rgb": the float's bits are the packed word
                    uint32_t bits;
                    memcpy(&bits, row.data() + props[i].offset, 4);
                    assign_packed(points[v], bits);
                } else {
                    assign(points[v], fields[i], read_binary(row.data() + props[i].offset, props[i].type));
                }
            }
        }
    } else {
        for (size_t k = 0; k < skip_before; k++)
            if (!fgets(line, sizeof(line), fp)) { ok = false; break; }
        for (size_t v = 0; v < nvertex && ok; v++) {
            if (!fgets(line, sizeof(line), fp)) { ok = false; break; }
            char *cur = line;
            for (size_t i = 0; i < props.size(); i++) {
                char *end = nullptr;
                const double val = strtod(cur, &end);
                if (end == cur) { ok = false; break; }
                cur = end;
                if (fields[i] == 7 && props[i].type == T_FLOAT32) {
                    const float f = (float)val;
                    uint32_t bits;
                    memcpy(&bits, &f, 4);
                    assign_packed(points[v], bits);
                } else {
                    assign(points[v], fields[i], val);
                }
            }
        }
    }
    fclose(fp);
    if (!ok) err = "truncated or malformed vertex data";
    return ok;
}

int write_ply(const char *filename, cwipc_pointcloud *pc, bool binary, const char *who) {
    if (pc == nullptr) {
        log(CWIPC_LOG_LEVEL_ERROR, who, "Saving NULL pointcloud not implemented");
        return -1;
    }
    const size_t bytes = pc->get_uncompressed_size();
    std::vector<cwipc_point> pts(bytes / sizeof(cwipc_point));
    if (bytes && pc->copy_uncompressed(pts.data(), bytes) < 0) {
        log(CWIPC_LOG_LEVEL_ERROR, who, std::string("Saving of PLY file failed: ") + filename);
        return -1;
    }
    FILE *fp = fopen(filename, "wb");
    if (!fp) {
        log(CWIPC_LOG_LEVEL_ERROR, who, std::string("Saving of PLY file failed: ") + filename);
        return -1;
    }
    fprintf(fp, "ply\nformat %s 1.0\ncomment cwipc_util_cuda generated\nelement vertex %zu\n", binary ? "binary_little_endian" : "ascii", pts.size());
    fprintf(fp, "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\nend_header\n");
    bool ok = true;
    if (binary) {
        // cwipc_point already is x,y,z float + r,g,b,alpha bytes: the file body is the raw array
        ok = pts.empty() || fwrite(pts.data(), sizeof(cwipc_point), pts.size(), fp) == pts.size();
    } else {
        for (const auto &p : pts) {
            if (fprintf(fp, "%.9g %.9g %.9g %u %u %u %u\n", p.x, p.y, p.z, p.r, p.g, p.b, p.tile) < 0) { ok = false; break; }
        }
    }
    if (fclose(fp) != 0) ok = false;
    if (!ok) {
        log(CWIPC_LOG_LEVEL_ERROR, who, std::string("Saving of PLY file failed: ") + filename);
        return -1;
    }
    return 0;
}

struct ErrorCapture {
    explicit ErrorCapture(char **errorMessage) { log_set_errorbuf(errorMessage); }
    ~ErrorCapture() { log_set_errorbuf(nullptr); }
};

} // namespace

extern "C" {

cwipc_pointcloud *cwipc_read(const char *filename, uint64_t timestamp, char **errorMessage, uint64_t apiVersion) {
    if (apiVersion < CWIPC_API_VERSION_OLD || apiVersion > CWIPC_API_VERSION) {
        if (errorMessage) {
            char *msg = (char *)malloc(1024);
            snprintf(msg, 1024, "cwipc_read: incorrect apiVersion 0x%08llx expected 0x%08llx..0x%08llx", (unsigned long long)apiVersion, (unsigned long long)CWIPC_API_VERSION_OLD,
                     (unsigned long long)CWIPC_API_VERSION);
            *errorMessage = msg;
        }
        return nullptr;
    }
    ErrorCapture cap(errorMessage);
    std::vector<cwipc_point> pts;
    std::string err;
    if (!read_ply(filename, pts, err)) {
        log(CWIPC_LOG_LEVEL_ERROR, "cwipc_read", std::string("Loading of PLY file failed: ") + filename + " (" + err + ")");
        return nullptr;
    }
    return guarded<cwipc_pointcloud *>("cwipc_read", nullptr, [&]() -> cwipc_pointcloud * { return DevicePointcloud::from_host(pts.data(), pts.size(), timestamp, true); });
}

int cwipc_write(const char *filename, cwipc_pointcloud *pc, char **errorMessage) {
    ErrorCapture cap(errorMessage);
    return write_ply(filename, pc, false, "cwipc_write");
}

int cwipc_write_ext(const char *filename, cwipc_pointcloud *pc, int flag, char **errorMessage) {
    ErrorCapture cap(errorMessage);
    return write_ply(filename, pc, (flag & CWIPC_FLAG_BINARY) != 0, "cwipc_write_ext");
}

} // extern "C"
