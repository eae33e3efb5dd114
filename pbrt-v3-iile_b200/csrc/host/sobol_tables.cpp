// The Sobol' generator matrices (data/sobol_matrices.bin, written by tools/extract_sobol_tables.py from
// the constants of src/core/sobolmatrices.cpp) linked into the library as a read-only blob.
#include <cstring>
#include <stdexcept>

#include "host_scene.h"

extern "C" {
extern const unsigned char iile_sobol_blob[];
extern const unsigned char iile_sobol_blob_end[];
}
// IILE_DATA_DIR is passed by the Makefile (-DIILE_DATA_DIR="..."): the assembler resolves .incbin there
asm(".section .rodata\n"
    ".balign 16\n"
    ".global iile_sobol_blob\n"
    "iile_sobol_blob:\n"
    ".incbin \"" IILE_DATA_DIR "/sobol_matrices.bin\"\n"
    ".global iile_sobol_blob_end\n"
    "iile_sobol_blob_end:\n"
    ".previous\n");

namespace iile {

const SobolTables &sobolTables() {
    static const SobolTables t = [] {
        SobolTables s{};
        const unsigned char *p = iile_sobol_blob;
        uint32_t head[8];
        std::memcpy(head, p, sizeof(head));
        if (std::memcmp(p, "SOBL", 4) != 0 || head[1] != 1024 || head[2] != 52)
            throw std::runtime_error("sobol_matrices.bin: bad header");
        s.nDims = (int)head[1];
        s.matrixSize = (int)head[2];
        s.nVdc = (int)head[3];
        s.nVdcInv = (int)head[4];
        s.matrices32 = reinterpret_cast<const uint32_t *>(p + 32);
        s.vdc = reinterpret_cast<const uint64_t *>(p + 32 + (size_t)s.nDims * s.matrixSize * 4);
        s.vdcInv = s.vdc + (size_t)s.nVdc * s.matrixSize;
        if ((const unsigned char *)(s.vdcInv + (size_t)s.nVdcInv * s.matrixSize) != iile_sobol_blob_end)
            throw std::runtime_error("sobol_matrices.bin: size does not match its header");
        return s;
    }();
    return t;
}

}  // namespace iile
