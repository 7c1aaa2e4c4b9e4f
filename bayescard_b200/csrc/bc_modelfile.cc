// Flat model file (".bcm"): the model a serving process needs, as ONE mmap-able file -- no pickle, no third-party
// classes (replaces `pickle.load` of the Bayescard_BN object,
// Models/BN_single_model.py:207-223, followed by init_inference_method, Models/Bayescard_BN.py:122-142).
//
//   header (128 bytes, little endian)
//     char     magic[8]      "BCB200M\0"
//     uint32   version       1
//     uint32   n_nodes
//     uint64   arena_floats, fan_floats, cpt64_doubles, meta_bytes
//     uint64   off_parent, off_card, off_cpt_off, off_stride, off_fan_off, off_arena, off_fan, off_cpt64, off_meta
//     uint64   file_bytes
//   sections (each 64 B aligned; the fp32 CPT arena 4096 B aligned so it can be handed to cudaMemcpy from the mapping)
//     parent int32[n] | card int32[n] | cpt_off int64[n] | stride int32[n] | fan_off int64[n]
//     arena fp32[arena_floats]      topologically ordered CPTs, rows 16 B aligned (the layout bc_model_create takes)
//     fan   fp32[fan_floats]
//     cpt64 fp64[...]               the same CPTs unpadded in fp64 (host-side checks / re-packing; not read here)
//     meta  utf-8 JSON              decode tables (encoding, n_in_bin, mapping, domain, ...) for the predicate compiler
// The Python writer / reader is bayescard_b200/loader.py (TreeModel.save_flat / load_flat).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "bc_internal.h"

namespace {
struct BcmHeader {
    char magic[8];
    uint32_t version, n_nodes;
    uint64_t arena_floats, fan_floats, cpt64_doubles, meta_bytes;
    uint64_t off_parent, off_card, off_cpt_off, off_stride, off_fan_off, off_arena, off_fan, off_cpt64, off_meta;
    uint64_t file_bytes;
};
static_assert(sizeof(BcmHeader) == 128, "header layout");
}  // namespace

extern "C" int bc_model_create_from_file(int device, const char* path, bc_model** out) {
    if (!path || !out) { bc_set_error("path/out is NULL"); return BC_EINVAL; }
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { bc_set_error("cannot open model file %s: %s", path, strerror(errno)); return BC_EINVAL; }
    struct stat st {};
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(BcmHeader)) {
        close(fd);
        bc_set_error("%s: not a flat model file (shorter than its header)", path);
        return BC_EINVAL;
    }
    const size_t bytes = (size_t)st.st_size;
    void* map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { bc_set_error("mmap of %s failed: %s", path, strerror(errno)); return BC_ENOMEM; }
    const unsigned char* base = static_cast<const unsigned char*>(map);
    BcmHeader h;
    std::memcpy(&h, base, sizeof(h));
    int rc = BC_OK;
    auto bad = [&](const char* why) {
        bc_set_error("%s: %s", path, why);
        rc = BC_EINVAL;
    };
    // counts are checked BEFORE they are multiplied: 4 * arena_floats wraps in uint64 for a crafted count
    auto inside = [&](uint64_t off, uint64_t count, uint64_t elem, uint64_t align) {
        return off % align == 0 && off <= bytes && count <= (bytes - off) / elem;
    };
    const uint64_t n = h.n_nodes;
    if (std::memcmp(h.magic, "BCB200M\0", 8) != 0) bad("bad magic (not a bayescard_b200 flat model file)");
    else if (h.version != 1) bad("unsupported flat model file version");
    else if (h.file_bytes != bytes) bad("truncated or padded file (size differs from the header)");
    else if (n == 0 || n > (1u << 20)) bad("implausible node count");
    else if (!inside(h.off_parent, n, 4, 4) || !inside(h.off_card, n, 4, 4) || !inside(h.off_cpt_off, n, 8, 8) ||
             !inside(h.off_stride, n, 4, 4) || !inside(h.off_fan_off, n, 8, 8) || !inside(h.off_arena, h.arena_floats, 4, 16) ||
             !inside(h.off_fan, h.fan_floats, 4, 16) || !inside(h.off_cpt64, h.cpt64_doubles, 8, 8) ||
             !inside(h.off_meta, h.meta_bytes, 1, 1))
        bad("a section lies outside the file");
    if (rc == BC_OK)
        rc = bc_model_create(device, (int)n, reinterpret_cast<const int32_t*>(base + h.off_parent),
                             reinterpret_cast<const int32_t*>(base + h.off_card), reinterpret_cast<const int64_t*>(base + h.off_cpt_off),
                             reinterpret_cast<const int32_t*>(base + h.off_stride), reinterpret_cast<const float*>(base + h.off_arena),
                             (size_t)h.arena_floats, reinterpret_cast<const int64_t*>(base + h.off_fan_off),
                             h.fan_floats ? reinterpret_cast<const float*>(base + h.off_fan) : nullptr, (size_t)h.fan_floats, out);
    munmap(map, bytes);
    return rc;
}
