// host_snapshot.cpp -- Gadget-2 (format 1) snapshot reader / writer of the host library (include/p2p_host.h).
//
// Replaces read_GadgetHeader / read_Particle_Gadget2 / write_Particle_Gadget2 (1_Indexing/src/snapshot.c:5-22,211-293,
// 397-503) for the device-resident stepping: a rank reads the slab [n_start, n_start + n_count) of the file's particle
// order straight into rows of doubles (what p2p_resident_load / p2p_route_load take) and writes its resident particles
// back.  Same conventions as the reference: positions and velocities float32 on disk, all particle types concatenated in
// type order, velocities multiplied by a^(3/2) on input and divided on output (its "gdt2unit"), everything written as
// type 1 (halo) particles with mass[1].  Unlike the reference the Fortran record markers are written with their proper
// values (the reference writes an uninitialised int and never checks them on input); ids are not stored, as there.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/p2p_host.h"

namespace {
#pragma pack(push, 1)
struct GadgetHeader {          // 1_Indexing/src/snapshot.c:5-22, 256 bytes
    int32_t npart[6];
    double mass[6];
    double time, redshift;
    int32_t flag_sfr, flag_feedback;
    int32_t npartTotal[6];
    int32_t flag_cooling, num_files;
    double BoxSize, Omega0, OmegaLambda, HubbleParam;
    char fill[256 - 6 * 4 - 6 * 8 - 2 * 8 - 2 * 4 - 6 * 4 - 2 * 4 - 4 * 8];
};
#pragma pack(pop)
static_assert(sizeof(GadgetHeader) == 256, "Gadget-2 header is 256 bytes");

struct File {
    FILE* f = nullptr;
    explicit File(const char* path, const char* mode) { f = fopen(path, mode); }
    ~File() { if (f) fclose(f); }
};

int read_header(FILE* f, GadgetHeader* h) {
    int32_t m0 = 0, m1 = 0;
    if (fread(&m0, 4, 1, f) != 1 || fread(h, sizeof *h, 1, f) != 1 || fread(&m1, 4, 1, f) != 1) return -1;
    return 0;
}

// block of 3 float32 per particle: rows [n_start, n_start + n_count) -> out (rows `stride` doubles apart), scaled
int read_block(FILE* f, int64_t ntot, int64_t n_start, int64_t n_count, double scale, double* out, int64_t stride) {
    int32_t marker = 0;
    if (fread(&marker, 4, 1, f) != 1) return -1;
    if (fseeko(f, (off_t)n_start * 12, SEEK_CUR) != 0) return -1;
    std::vector<float> buf((size_t)(1 << 20) * 3);
    for (int64_t done = 0; done < n_count;) {
        const int64_t n = n_count - done < (1 << 20) ? n_count - done : (1 << 20);
        if (fread(buf.data(), 12, (size_t)n, f) != (size_t)n) return -1;
        if (out)
            for (int64_t i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) out[(done + i) * stride + k] = (double)buf[(size_t)(3 * i + k)] * scale;
        done += n;
    }
    if (fseeko(f, (off_t)(ntot - n_start - n_count) * 12, SEEK_CUR) != 0) return -1;
    if (fread(&marker, 4, 1, f) != 1) return -1;
    return 0;
}
}  // namespace

extern "C" {

int p2p_snapshot_header(const char* path, p2p_snapshot_info* info) {
    if (!path || !info) return -2;
    File F(path, "rb");
    if (!F.f) return -1;
    GadgetHeader h;
    if (read_header(F.f, &h)) return -1;
    memset(info, 0, sizeof *info);
    for (int k = 0; k < 6; k++) { info->npart[k] = h.npart[k]; info->npart_total[k] = (uint32_t)h.npartTotal[k]; info->mass[k] = h.mass[k]; info->nfile += h.npart[k]; }
    info->time = h.time; info->redshift = h.redshift; info->box = h.BoxSize; info->omega0 = h.Omega0; info->omega_lambda = h.OmegaLambda;
    info->hubble = h.HubbleParam; info->num_files = h.num_files;
    return 0;
}

int p2p_snapshot_read(const char* path, int64_t n_start, int64_t n_count, double* pos, int64_t pos_stride, double* vel, int64_t vel_stride) {
    if (!path || n_start < 0 || n_count < 0 || (pos && pos_stride < 3) || (vel && vel_stride < 3)) return -2;
    File F(path, "rb");
    if (!F.f) return -1;
    GadgetHeader h;
    if (read_header(F.f, &h)) return -1;
    int64_t ntot = 0;
    for (int k = 0; k < 6; k++) ntot += h.npart[k];
    if (n_start + n_count > ntot) return -2;
    if (read_block(F.f, ntot, n_start, n_count, 1.0, pos, pos_stride)) return -1;
    const double gdt2unit = pow(1.0 / (1.0 + h.redshift), 1.5);          // 1_Indexing/src/snapshot.c:262
    if (vel && read_block(F.f, ntot, n_start, n_count, gdt2unit, vel, vel_stride)) return -1;
    return 0;
}

int p2p_snapshot_write(const char* path, const p2p_snapshot_info* info, int64_t n_count, const double* pos, int64_t pos_stride, const double* vel,
                       int64_t vel_stride) {
    if (!path || !info || n_count < 0 || n_count > 0x7fffffffLL || (n_count && (!pos || pos_stride < 3)) || (vel && vel_stride < 3)) return -2;
    File F(path, "wb");
    if (!F.f) return -1;
    GadgetHeader h;
    memset(&h, 0, sizeof h);
    h.npart[1] = (int32_t)n_count;                                       // everything is written as type 1, as the reference does
    for (int k = 0; k < 6; k++) h.mass[k] = info->mass[k];
    h.npartTotal[1] = (int32_t)(info->npart_total[1] ? info->npart_total[1] : (uint32_t)n_count);
    h.num_files = 1;
    h.BoxSize = info->box; h.Omega0 = info->omega0; h.OmegaLambda = info->omega_lambda; h.HubbleParam = info->hubble;
    h.time = 1.0 / (info->redshift + 1.0); h.redshift = info->redshift;
    const int32_t m256 = 256, mblk = (int32_t)(n_count * 12);
    if (fwrite(&m256, 4, 1, F.f) != 1 || fwrite(&h, sizeof h, 1, F.f) != 1 || fwrite(&m256, 4, 1, F.f) != 1) return -1;
    const double gdt2unit = pow(1.0 / (1.0 + h.redshift), 1.5);
    std::vector<float> buf((size_t)(1 << 20) * 3);
    for (int blk = 0; blk < 2; blk++) {
        const double* src = blk ? vel : pos;
        const int64_t stride = blk ? vel_stride : pos_stride;
        if (fwrite(&mblk, 4, 1, F.f) != 1) return -1;
        for (int64_t done = 0; done < n_count;) {
            const int64_t n = n_count - done < (1 << 20) ? n_count - done : (1 << 20);
            for (int64_t i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) {
                    // (float)vel / gdt2unit: the reference narrows first, then divides in double and narrows again
                    const double v = src ? src[(done + i) * stride + k] : 0.0;
                    buf[(size_t)(3 * i + k)] = blk ? (float)((double)(float)v / gdt2unit) : (float)v;
                }
            if (fwrite(buf.data(), 12, (size_t)n, F.f) != (size_t)n) return -1;
            done += n;
        }
        if (fwrite(&mblk, 4, 1, F.f) != 1) return -1;
    }
    return 0;
}

}  // extern "C"
