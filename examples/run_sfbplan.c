/*
 * run_sfbplan.c -- a stencil program executed from plain C through the per-program handle of
 * include/sfb200.h (sfb_program_*), no Python in the process.
 *
 * The counterpart in the reference is the ctypes sequence of dace/dace/codegen/compiled_sdfg.py:
 * load the program's library (:20-150), __dace_init (:182-185), __program(handle, arrays...) (:286-294),
 * __dace_exit (:256-267).  Here the "library" is a compiled image plus a plan script that
 * CudaProgram.export_plan() writes next to it:
 *
 *     image  <file.cubin>
 *     buffer <field> <bytes> <share_with | -1>
 *     launch <kernel> <gx> <gy> <gz> <bx> <by> <bz> <dynamic_smem> <num_params>
 *       bytes  <n> <hex digits>
 *       buffer <buffer index> <byte offset>
 *       tmap   <buffer index> <dtype> <rank> <dims...> <strides_bytes (rank-1)...> <box...>
 *       table  <n words> <int32 words...>
 *     input  <field> <file.dat>          raw little-endian array, as the reference's .dat inputs
 *     output <field> <file.dat>
 *
 * usage: run_sfbplan <plan.sfbplan> [device] [repetitions]
 * build: gcc -O2 -I include examples/run_sfbplan.c -o run_sfbplan -L stencilflow_b200 -lsfb200 \
 *            -Wl,-rpath,$PWD/stencilflow_b200          (libsfb200.so is built by stencilflow_b200/build.py)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sfb200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ < 0) {                                                                \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sfb_last_error());    \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static void* read_file(const char* path, size_t* size) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* buf = malloc(n > 0 ? (size_t)n : 1);
    if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    if (size) *size = (size_t)n;
    return buf;
}

static int hex_nibble(int c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }

struct io { char field[128]; char path[1024]; int is_output; void* host; size_t bytes; };

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s plan.sfbplan [device] [repetitions]\n", argv[0]);
        return 2;
    }
    int device = argc > 2 ? atoi(argv[2]) : 0, reps = argc > 3 ? atoi(argv[3]) : 0;
    FILE* plan = fopen(argv[1], "r");
    if (!plan) { perror(argv[1]); return 2; }
    CHECK(sfb_init(device));
    sfb_program* prog = NULL;
    struct io ios[64];
    int n_io = 0;
    char word[64], name[1024];
    while (fscanf(plan, "%63s", word) == 1) {
        if (!strcmp(word, "image")) {
            if (fscanf(plan, "%1023s", name) != 1) return 2;
            size_t size = 0;
            void* image = read_file(name, &size);
            if (!image) { perror(name); return 2; }
            CHECK(sfb_program_create(image, size, &prog));
            free(image);
        } else if (!strcmp(word, "buffer")) {
            unsigned long long bytes; int share, index;
            if (fscanf(plan, "%1023s %llu %d", name, &bytes, &share) != 3) return 2;
            CHECK(sfb_program_add_buffer(prog, name, (size_t)bytes, share, &index));
        } else if (!strcmp(word, "launch")) {
            unsigned grid[3], block[3], smem; int np;
            if (fscanf(plan, "%1023s %u %u %u %u %u %u %u %d", name, &grid[0], &grid[1], &grid[2], &block[0], &block[1],
                       &block[2], &smem, &np) != 9) return 2;
            sfb_launch_param* ps = (sfb_launch_param*)calloc((size_t)(np > 0 ? np : 1), sizeof(sfb_launch_param));
            void** owned = (void**)calloc((size_t)(np > 0 ? np : 1), sizeof(void*));
            for (int k = 0; k < np; ++k) {
                char kind[16];
                if (fscanf(plan, "%15s", kind) != 1) return 2;
                if (!strcmp(kind, "bytes")) {
                    unsigned n; char hex[256];
                    if (fscanf(plan, "%u %255s", &n, hex) != 2 || strlen(hex) != 2 * n) return 2;
                    unsigned char* v = (unsigned char*)malloc(n);
                    for (unsigned b = 0; b < n; ++b) v[b] = (unsigned char)(hex_nibble(hex[2 * b]) * 16 + hex_nibble(hex[2 * b + 1]));
                    ps[k].kind = SFB_PARAM_BYTES; ps[k].data = v; ps[k].size = n; owned[k] = v;
                } else if (!strcmp(kind, "buffer")) {
                    int b; unsigned long long off;
                    if (fscanf(plan, "%d %llu", &b, &off) != 2) return 2;
                    ps[k].kind = SFB_PARAM_BUFFER; ps[k].buffer = b; ps[k].offset = off;
                } else if (!strcmp(kind, "tmap")) {
                    int b, dt, rank;
                    if (fscanf(plan, "%d %d %d", &b, &dt, &rank) != 3 || rank < 1 || rank > 5) return 2;
                    ps[k].kind = SFB_PARAM_TMAP; ps[k].buffer = b; ps[k].dtype = dt; ps[k].rank = rank;
                    unsigned long long v;
                    for (int d = 0; d < rank; ++d) { if (fscanf(plan, "%llu", &v) != 1) return 2; ps[k].dims[d] = v; }
                    for (int d = 0; d + 1 < rank; ++d) { if (fscanf(plan, "%llu", &v) != 1) return 2; ps[k].strides_bytes[d] = v; }
                    for (int d = 0; d < rank; ++d) { if (fscanf(plan, "%llu", &v) != 1) return 2; ps[k].box[d] = (uint32_t)v; }
                } else if (!strcmp(kind, "table")) {
                    unsigned n;
                    if (fscanf(plan, "%u", &n) != 1) return 2;
                    int32_t* t = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
                    for (unsigned w = 0; w < n; ++w) { int x; if (fscanf(plan, "%d", &x) != 1) return 2; t[w] = x; }
                    ps[k].kind = SFB_PARAM_TABLE; ps[k].data = t; ps[k].size = (uint32_t)(n * sizeof(int32_t)); owned[k] = t;
                } else {
                    fprintf(stderr, "unknown parameter kind %s\n", kind);
                    return 2;
                }
            }
            CHECK(sfb_program_add_launch(prog, name, grid, block, smem, np, ps));
            for (int k = 0; k < np; ++k) free(owned[k]);
            free(owned);
            free(ps);
        } else if (!strcmp(word, "input") || !strcmp(word, "output")) {
            struct io* x = &ios[n_io++];
            x->is_output = word[0] == 'o';
            if (fscanf(plan, "%127s %1023s", x->field, x->path) != 2) return 2;
            size_t dbytes = 0;
            CHECK(sfb_program_buffer(prog, x->field, NULL, &dbytes));
            if (x->is_output) {
                x->bytes = dbytes;
                x->host = calloc(1, dbytes ? dbytes : 1);
            } else {
                x->host = read_file(x->path, &x->bytes);
                if (!x->host) { perror(x->path); return 2; }
            }
            CHECK(sfb_program_bind(prog, x->field, x->host, x->bytes, x->is_output));
        } else {
            fprintf(stderr, "unknown directive %s\n", word);
            return 2;
        }
    }
    fclose(plan);
    if (!prog) { fprintf(stderr, "plan has no image\n"); return 2; }
    int n_launch = 0;
    CHECK(sfb_program_num_launches(prog, &n_launch));
    CHECK(sfb_program_call(prog, NULL));               /* inputs in, every launch, outputs back */
    for (int k = 0; k < n_io; ++k) {
        if (!ios[k].is_output) continue;
        FILE* f = fopen(ios[k].path, "wb");
        if (!f || fwrite(ios[k].host, 1, ios[k].bytes, f) != ios[k].bytes) { perror(ios[k].path); return 1; }
        fclose(f);
    }
    printf("ran %d launch(es)", n_launch);
    if (reps > 0) {
        float ms = 0.0f;
        CHECK(sfb_program_run(prog, reps, NULL, &ms));
        printf(", %d repetitions in %.3f ms on the device", reps, ms);
    }
    printf("\n");
    CHECK(sfb_program_destroy(prog));
    for (int k = 0; k < n_io; ++k) free(ios[k].host);
    return 0;
}
