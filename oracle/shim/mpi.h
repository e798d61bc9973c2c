// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Minimal single-node stand-in for the 11 MPI entry points the reference uses, so that the UNMODIFIED
// sources under /root/reference compile into oracle/_ref/ in an image that has no MPI at all.
// Ranks are fork()ed children of rank 0 that share one anonymous mapping:
//   [Shared header | NP mailboxes of SLOT bytes | bump arena for "shared windows"].
// Rank count comes from the environment variable PIMDB_NP (one rank per bead, as the reference requires:
// /root/reference/src/simulation.cpp:19).
//
// Call sites served (reference file:line):
//   MPI_Sendrecv ring shift ............ src/simulation.cpp:302-346
//   MPI_Allreduce(SUM, double) ......... src/simulation.cpp:595, src/observables/observable.cpp:105
//   MPI_Barrier / MPI_Wtime ............ src/simulation.cpp:225,282 and the normal-mode code
//   MPI_Win_allocate_shared / query .... src/normal_modes.cpp:13-46
// Every collective in the reference is entered by all ranks in lock-step, which is what makes the
// "publish to own mailbox, barrier, read peer mailbox, barrier" scheme below valid.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct { size_t off; } MPI_Win;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8 /* doubles as the element size in bytes */
#define MPI_SUM 0
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

namespace mpishim {
struct Shared {
    pthread_barrier_t bar;
    size_t arena_used;
    size_t last_win_off;
};
constexpr size_t kSlot = size_t(1) << 22;   // 4 MiB mailbox per rank
constexpr size_t kArena = size_t(1) << 31;  // 2 GiB of lazily-backed window space
inline Shared* sh = nullptr;
inline char* slots = nullptr;
inline char* arena = nullptr;
inline int rank = 0, np = 1;
inline char* slot_of(int r) { return slots + kSlot * size_t(r); }
}  // namespace mpishim

inline int MPI_Barrier(MPI_Comm) {
    pthread_barrier_wait(&mpishim::sh->bar);
    return 0;
}

inline int MPI_Init(int*, char***) {
    using namespace mpishim;
    const char* e = getenv("PIMDB_NP");
    np = e ? atoi(e) : 1;
    if (np < 1) np = 1;
    size_t total = 4096 + kSlot * size_t(np) + kArena;
    char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE,
                             MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == MAP_FAILED) { perror("mpishim mmap"); exit(1); }
    sh = (Shared*)base;
    slots = base + 4096;
    arena = slots + kSlot * size_t(np);
    pthread_barrierattr_t attr;
    pthread_barrierattr_init(&attr);
    pthread_barrierattr_setpshared(&attr, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &attr, (unsigned)np);
    sh->arena_used = 0;
    fflush(stdout);
    fflush(stderr);
    for (int r = 1; r < np; ++r) {
        pid_t pid = fork();
        if (pid < 0) { perror("mpishim fork"); exit(1); }
        if (pid == 0) { rank = r; break; }
    }
    return 0;
}

inline int MPI_Finalize() {
    fflush(stdout);
    fflush(stderr);
    MPI_Barrier(0);
    if (mpishim::rank != 0) _exit(0);
    while (wait(nullptr) > 0) {}
    return 0;
}

inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = mpishim::rank; return 0; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = mpishim::np; return 0; }

inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline int MPI_Sendrecv(const void* sbuf, int scount, MPI_Datatype stype, int /*dest*/, int /*stag*/,
                        void* rbuf, int rcount, MPI_Datatype rtype, int source, int /*rtag*/,
                        MPI_Comm, MPI_Status*) {
    using namespace mpishim;
    memcpy(slot_of(rank), sbuf, size_t(scount) * size_t(stype));
    MPI_Barrier(0);
    memcpy(rbuf, slot_of(source), size_t(rcount) * size_t(rtype));
    MPI_Barrier(0);
    return 0;
}

inline int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
    using namespace mpishim;
    memcpy(slot_of(rank), sbuf, size_t(count) * size_t(type));
    MPI_Barrier(0);
    double* out = (double*)rbuf;
    for (int i = 0; i < count; ++i) {
        double acc = 0.0;
        for (int r = 0; r < np; ++r) acc += ((const double*)slot_of(r))[i];
        out[i] = acc;
    }
    MPI_Barrier(0);
    return 0;
}

inline int MPI_Win_allocate_shared(MPI_Aint bytes, int, MPI_Info, MPI_Comm, void* baseptr, MPI_Win* win) {
    using namespace mpishim;
    MPI_Barrier(0);
    if (rank == 0) {
        sh->last_win_off = sh->arena_used;
        sh->arena_used += (size_t(bytes) + 63) & ~size_t(63);
        if (sh->arena_used > kArena) { fprintf(stderr, "mpishim: window arena exhausted\n"); abort(); }
    }
    MPI_Barrier(0);
    win->off = sh->last_win_off;
    *(void**)baseptr = arena + win->off;
    return 0;
}

inline int MPI_Win_shared_query(MPI_Win win, int, MPI_Aint*, int*, void* baseptr) {
    *(void**)baseptr = mpishim::arena + win.off;
    return 0;
}

inline int MPI_Win_free(MPI_Win*) { return 0; }
