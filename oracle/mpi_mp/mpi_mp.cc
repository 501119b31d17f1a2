// oracle/mpi_mp/mpi_mp.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Runtime of the multi-process MPI replacement declared in oracle/mpi_mp/mpi.h.  Started by oracle/mprun.py, which
// creates one Unix socket pair per pair of ranks and passes every rank its rank, the world size and the descriptors
// (SB200_MPI_RANK, SB200_MPI_SIZE, SB200_MPI_FDS = descriptor towards rank 0, 1, ... with -1 for itself).
//
//   * point to point: a send writes {context, tag, bytes} + the packed payload to the peer's socket under a per-peer
//     lock; a progress thread per rank reads every socket into per-source queues, so a send never waits for the matching
//     receive; a receive takes the first queued message of its source with its (context, tag).  Non-blocking receives
//     complete in MPI_Wait / MPI_Waitall (the data is buffered by then or arrives while waiting).
//   * communicators: a context id + the list of world ranks.  MPI_Comm_create_group derives the new context from the
//     parent's, the member list, the tag and the number of times this process created that combination -- the same on
//     every member without any message.
//   * collectives: linear algorithms over point-to-point messages with negative tags numbered per communicator;
//     reductions combine the contributions in rank order on rank 0 of the communicator (deterministic).
//   * datatypes: basic codes, contiguous and vector types (packed / unpacked around the transfer).
#include "mpi.h"

#include <poll.h>
#include <pthread.h>
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Msg { int ctx, tag; std::vector<char> data; };
struct Comm { int ctx = 0; std::vector<int> ranks; int me = -1; bool used = false; long coll_seq = 0; };
struct DType { int count, blocklen, stride; MPI_Datatype base; bool used; };
struct Req { bool used = false, recv = false; void* buf = nullptr; int n = 0; MPI_Datatype t = 0; int src_world = 0, ctx = 0, tag = 0; };

int g_rank = 0, g_size = 1;
bool g_init = false;
std::vector<int> g_fd;
std::vector<std::mutex*> g_send_mu;
std::mutex g_in_mu;
std::condition_variable g_in_cv;
std::vector<std::deque<Msg>> g_inbox;
std::thread g_progress;
volatile bool g_stop = false;

std::mutex g_tab_mu;                      // communicators, groups, datatypes, requests, ops
std::vector<Comm> g_comms(3);
std::vector<std::vector<int>> g_groups(1);
std::vector<DType> g_types;
std::vector<Req> g_reqs(1);
std::vector<MPI_User_function*> g_ops;
std::map<std::string, int> g_create_seq;

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "mpi_mp[rank %d]: %s (errno %d: %s)\n", g_rank, what, errno, strerror(errno));
    abort();
}

void write_all(int fd, const void* p, size_t n)
{
    const char* c = static_cast<const char*>(p);
    while (n > 0) {
        ssize_t w = write(fd, c, n);
        if (w < 0) { if (errno == EINTR) continue; die("write to a peer failed"); }
        c += w; n -= size_t(w);
    }
}

bool read_all(int fd, void* p, size_t n)
{
    char* c = static_cast<char*>(p);
    while (n > 0) {
        ssize_t r = read(fd, c, n);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        c += r; n -= size_t(r);
    }
    return true;
}

void deliver(int src, Msg&& m)
{
    { std::lock_guard<std::mutex> lk(g_in_mu); g_inbox[size_t(src)].push_back(std::move(m)); }
    g_in_cv.notify_all();
}

void progress_loop()
{
    std::vector<pollfd> pf;
    std::vector<int> who;
    for (int r = 0; r < g_size; ++r)
        if (r != g_rank) { pf.push_back({g_fd[size_t(r)], POLLIN, 0}); who.push_back(r); }
    while (! g_stop) {
        if (pf.empty()) { usleep(20000); continue; }
        int k = poll(pf.data(), nfds_t(pf.size()), 50);
        if (k <= 0) continue;
        for (size_t i = 0; i < pf.size(); ++i) {
            if (! (pf[i].revents & (POLLIN | POLLHUP))) continue;
            int64_t hdr[3];
            if (! read_all(pf[i].fd, hdr, sizeof hdr)) { pf[i].fd = -1; continue; }      // peer gone: stop polling it
            Msg m;
            m.ctx = int(hdr[0]); m.tag = int(hdr[1]);
            m.data.resize(size_t(hdr[2]));
            if (hdr[2] > 0 && ! read_all(pf[i].fd, m.data.data(), size_t(hdr[2]))) { pf[i].fd = -1; continue; }
            deliver(who[i], std::move(m));
        }
    }
}

void raw_send(int dst_world, int ctx, int tag, const void* data, size_t bytes)
{
    if (dst_world == g_rank) {
        Msg m; m.ctx = ctx; m.tag = tag; m.data.assign(static_cast<const char*>(data), static_cast<const char*>(data) + bytes);
        deliver(g_rank, std::move(m));
        return;
    }
    int64_t hdr[3] = {ctx, tag, int64_t(bytes)};
    std::lock_guard<std::mutex> lk(*g_send_mu[size_t(dst_world)]);
    write_all(g_fd[size_t(dst_world)], hdr, sizeof hdr);
    if (bytes) write_all(g_fd[size_t(dst_world)], data, bytes);
}

std::vector<char> raw_recv(int src_world, int ctx, int tag)
{
    std::unique_lock<std::mutex> lk(g_in_mu);
    for (;;) {
        auto& q = g_inbox[size_t(src_world)];
        for (auto it = q.begin(); it != q.end(); ++it)
            if (it->ctx == ctx && it->tag == tag) {
                std::vector<char> d = std::move(it->data);
                q.erase(it);
                return d;
            }
        g_in_cv.wait(lk);
    }
}

// ---- datatypes
const int BASIC_SIZE[12] = {0, 1, 4, 4, 8, 4, 8, 8, 16, 8, 8, 16};

DType derived(MPI_Datatype t)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (t < 1000 || size_t(t - 1000) >= g_types.size() || ! g_types[size_t(t - 1000)].used) die("unknown datatype");
    return g_types[size_t(t - 1000)];
}
size_t type_size(MPI_Datatype t)
{
    if (t >= 1 && t <= 11) return size_t(BASIC_SIZE[t]);
    DType d = derived(t);
    return size_t(d.count) * d.blocklen * type_size(d.base);
}
size_t type_extent(MPI_Datatype t)
{
    if (t >= 1 && t <= 11) return size_t(BASIC_SIZE[t]);
    DType d = derived(t);
    return (size_t(d.count - 1) * d.stride + d.blocklen) * type_extent(d.base);
}
void pack_one(const char* src, MPI_Datatype t, std::vector<char>& out)
{
    if (t >= 1 && t <= 11) { out.insert(out.end(), src, src + BASIC_SIZE[t]); return; }
    DType d = derived(t);
    const size_t be = type_extent(d.base);
    for (int c = 0; c < d.count; ++c)
        for (int b = 0; b < d.blocklen; ++b) pack_one(src + (size_t(c) * d.stride + b) * be, d.base, out);
}
const char* unpack_one(const char* in, char* dst, MPI_Datatype t)
{
    if (t >= 1 && t <= 11) { memcpy(dst, in, size_t(BASIC_SIZE[t])); return in + BASIC_SIZE[t]; }
    DType d = derived(t);
    const size_t be = type_extent(d.base);
    for (int c = 0; c < d.count; ++c)
        for (int b = 0; b < d.blocklen; ++b) in = unpack_one(in, dst + (size_t(c) * d.stride + b) * be, d.base);
    return in;
}
std::vector<char> pack(const void* buf, int n, MPI_Datatype t)
{
    std::vector<char> out;
    if (t >= 1 && t <= 11) { out.assign(static_cast<const char*>(buf), static_cast<const char*>(buf) + size_t(n) * BASIC_SIZE[t]); return out; }
    out.reserve(size_t(n) * type_size(t));
    const size_t ext = type_extent(t);
    for (int i = 0; i < n; ++i) pack_one(static_cast<const char*>(buf) + size_t(i) * ext, t, out);
    return out;
}
void unpack(const std::vector<char>& in, void* buf, int n, MPI_Datatype t)
{
    const size_t want = size_t(n) * type_size(t);
    if (in.size() > want) die("received message longer than the receive buffer");
    if (t >= 1 && t <= 11) { memcpy(buf, in.data(), in.size()); return; }
    const size_t ext = type_extent(t), one = type_size(t);
    const char* p = in.data();
    for (int i = 0; i < n && size_t(p - in.data()) + one <= in.size(); ++i) p = unpack_one(p, static_cast<char*>(buf) + size_t(i) * ext, t);
}

Comm get_comm(MPI_Comm c)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (c <= 0 || size_t(c) >= g_comms.size() || ! g_comms[size_t(c)].used) die("invalid communicator");
    return g_comms[size_t(c)];
}
int next_coll_tag(MPI_Comm c)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    const long s = g_comms[size_t(c)].coll_seq++;
    return -int(2 + (s & 0x3fffffff));
}

// ---- reductions
template <typename T> void red_basic(int op, const T* in, T* io, int n)
{
    for (int i = 0; i < n; ++i)
        switch (op) {
            case MPI_SUM:  io[i] = io[i] + in[i]; break;
            case MPI_PROD: io[i] = io[i] * in[i]; break;
            case MPI_MAX:  io[i] = in[i] > io[i] ? in[i] : io[i]; break;
            case MPI_MIN:  io[i] = in[i] < io[i] ? in[i] : io[i]; break;
            case MPI_LAND: io[i] = T(io[i] && in[i]); break;
            case MPI_LOR:  io[i] = T(io[i] || in[i]); break;
            default: die("reduction operation not implemented for this type");
        }
}
template <typename V> void red_loc(int op, const void* in_, void* io_, int n)
{
    struct P { V v; int loc; };
    const P* in = static_cast<const P*>(in_);
    P* io = static_cast<P*>(io_);
    for (int i = 0; i < n; ++i) {
        const bool take = op == MPI_MAXLOC ? (in[i].v > io[i].v || (in[i].v == io[i].v && in[i].loc < io[i].loc))
                                           : (in[i].v < io[i].v || (in[i].v == io[i].v && in[i].loc < io[i].loc));
        if (take) io[i] = in[i];
    }
}
// io = in (op) io, `in` being the contribution of the LOWER rank (the order MPI prescribes for non-commutative ops)
void reduce_local(MPI_Op op, MPI_Datatype t, const void* in, void* io, int n)
{
    if (op >= 100) {
        MPI_User_function* f;
        { std::lock_guard<std::mutex> lk(g_tab_mu); f = g_ops[size_t(op - 100)]; }
        f(const_cast<void*>(in), io, &n, &t);
        return;
    }
    if (op == MPI_MAXLOC || op == MPI_MINLOC) {
        if (t == MPI_DOUBLE_INT) red_loc<double>(op, in, io, n);
        else if (t == MPI_FLOAT_INT) red_loc<float>(op, in, io, n);
        else if (t == MPI_2INT) red_loc<int>(op, in, io, n);
        else die("MAXLOC / MINLOC on an unsupported type");
        return;
    }
    switch (t) {
        case 1: red_basic(op, static_cast<const signed char*>(in), static_cast<signed char*>(io), n); break;
        case 2: red_basic(op, static_cast<const int*>(in), static_cast<int*>(io), n); break;
        case 3: red_basic(op, static_cast<const unsigned*>(in), static_cast<unsigned*>(io), n); break;
        case 4: red_basic(op, static_cast<const long*>(in), static_cast<long*>(io), n); break;
        case 5: red_basic(op, static_cast<const float*>(in), static_cast<float*>(io), n); break;
        case 6: red_basic(op, static_cast<const double*>(in), static_cast<double*>(io), n); break;
        case 7: if (op != MPI_SUM) die("complex reduction other than SUM"); red_basic(op, static_cast<const float*>(in), static_cast<float*>(io), 2 * n); break;
        case 8: if (op != MPI_SUM) die("complex reduction other than SUM"); red_basic(op, static_cast<const double*>(in), static_cast<double*>(io), 2 * n); break;
        default: die("reduction on an unsupported datatype");
    }
}

int new_request(const Req& r)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    for (size_t i = 1; i < g_reqs.size(); ++i)
        if (! g_reqs[i].used) { g_reqs[i] = r; g_reqs[i].used = true; return int(i); }
    g_reqs.push_back(r);
    g_reqs.back().used = true;
    return int(g_reqs.size() - 1);
}

} // namespace

extern "C" {

int MPI_Init_thread(int* argc, char*** argv, int required, int* provided)
{
    (void) argc; (void) argv;
    if (provided) *provided = required;
    if (g_init) return MPI_SUCCESS;
    const char* r = getenv("SB200_MPI_RANK");
    const char* s = getenv("SB200_MPI_SIZE");
    const char* f = getenv("SB200_MPI_FDS");
    g_rank = r ? atoi(r) : 0;
    g_size = s ? atoi(s) : 1;
    g_fd.assign(size_t(g_size), -1);
    if (g_size > 1) {
        if (! f) die("SB200_MPI_FDS missing: start the program through oracle/mprun.py");
        std::string fs(f);
        size_t pos = 0;
        for (int i = 0; i < g_size; ++i) {
            size_t c = fs.find(',', pos);
            g_fd[size_t(i)] = atoi(fs.substr(pos, c == std::string::npos ? std::string::npos : c - pos).c_str());
            pos = c == std::string::npos ? fs.size() : c + 1;
        }
    }
    g_inbox.resize(size_t(g_size));
    for (int i = 0; i < g_size; ++i) g_send_mu.push_back(new std::mutex);
    g_comms[1].used = true; g_comms[1].ctx = 1; g_comms[1].me = g_rank;
    for (int i = 0; i < g_size; ++i) g_comms[1].ranks.push_back(i);
    g_comms[2].used = true; g_comms[2].ctx = 2; g_comms[2].me = 0; g_comms[2].ranks = {g_rank};
    g_init = true;
    if (g_size > 1) g_progress = std::thread(progress_loop);
    return MPI_SUCCESS;
}
int MPI_Init(int* argc, char*** argv) { int p; return MPI_Init_thread(argc, argv, MPI_THREAD_MULTIPLE, &p); }
int MPI_Initialized(int* flag) { *flag = g_init ? 1 : 0; return MPI_SUCCESS; }
int MPI_Finalize(void)
{
    if (! g_init) return MPI_SUCCESS;
    MPI_Barrier(MPI_COMM_WORLD);
    g_stop = true;
    if (g_progress.joinable()) g_progress.join();
    return MPI_SUCCESS;
}
double MPI_Wtime(void)
{
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return double(t.tv_sec) + 1e-9 * double(t.tv_nsec);
}
int MPI_Error_string(int code, char* str, int* len) { *len = snprintf(str, MPI_MAX_ERROR_STRING, "mpi_mp error %d", code); return MPI_SUCCESS; }

int MPI_Comm_rank(MPI_Comm c, int* rank) { *rank = get_comm(c).me; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* size) { *size = int(get_comm(c).ranks.size()); return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g)
{
    Comm cm = get_comm(c);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    g_groups.push_back(cm.ranks);
    *g = int(g_groups.size() - 1);
    return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* c)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (*c > 2 && size_t(*c) < g_comms.size()) g_comms[size_t(*c)].used = false;
    *c = MPI_COMM_NULL;
    return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* out)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (g <= 0 || size_t(g) >= g_groups.size()) die("invalid group");
    std::vector<int> v;
    for (int i = 0; i < n; ++i) v.push_back(g_groups[size_t(g)][size_t(ranks[i])]);
    g_groups.push_back(v);
    *out = int(g_groups.size() - 1);
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group* g) { *g = MPI_GROUP_NULL; return MPI_SUCCESS; }       // group tables are tiny: never reclaimed
int MPI_Group_translate_ranks(MPI_Group a, int n, const int* in, MPI_Group b, int* out)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    const auto& ga = g_groups[size_t(a)];
    const auto& gb = g_groups[size_t(b)];
    for (int i = 0; i < n; ++i) {
        out[i] = MPI_UNDEFINED;
        const int w = ga[size_t(in[i])];
        for (size_t j = 0; j < gb.size(); ++j) if (gb[j] == w) { out[i] = int(j); break; }
    }
    return MPI_SUCCESS;
}
int MPI_Comm_create_group(MPI_Comm c, MPI_Group g, int tag, MPI_Comm* out)
{
    Comm parent = get_comm(c);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    const std::vector<int>& members = g_groups[size_t(g)];
    int me = -1;
    for (size_t i = 0; i < members.size(); ++i) if (members[i] == g_rank) me = int(i);
    if (me < 0) { *out = MPI_COMM_NULL; return MPI_SUCCESS; }
    std::string key = std::to_string(parent.ctx) + "/" + std::to_string(tag) + ":";
    for (int r : members) key += std::to_string(r) + ",";
    const int seq = g_create_seq[key]++;
    // FNV-1a over the key and the sequence number: the same on every member, distinct from the fixed contexts 1 and 2
    uint32_t h = 2166136261u;
    for (char ch : key + "#" + std::to_string(seq)) { h ^= uint8_t(ch); h *= 16777619u; }
    Comm n;
    n.used = true; n.ctx = int(h & 0x7fffffff) | 0x100; n.ranks = members; n.me = me;
    size_t slot = 0;
    for (size_t i = 3; i < g_comms.size(); ++i) if (! g_comms[i].used) { slot = i; break; }
    if (slot == 0) { g_comms.push_back(n); slot = g_comms.size() - 1; } else g_comms[slot] = n;
    *out = int(slot);
    return MPI_SUCCESS;
}

int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype* t)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    DType d{count, blocklen, stride, old, true};
    for (size_t i = 0; i < g_types.size(); ++i)
        if (! g_types[i].used) { g_types[i] = d; *t = MPI_Datatype(1000 + i); return MPI_SUCCESS; }
    g_types.push_back(d);
    *t = MPI_Datatype(1000 + g_types.size() - 1);
    return MPI_SUCCESS;
}
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype* t) { return MPI_Type_vector(1, count, count, old, t); }
int MPI_Type_commit(MPI_Datatype* t) { (void) t; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* t)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    if (*t >= 1000 && size_t(*t - 1000) < g_types.size()) g_types[size_t(*t - 1000)].used = false;
    *t = 0;
    return MPI_SUCCESS;
}
int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op)
{
    (void) commute;
    std::lock_guard<std::mutex> lk(g_tab_mu);
    g_ops.push_back(f);
    *op = MPI_Op(100 + g_ops.size() - 1);
    return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op* op) { *op = 0; return MPI_SUCCESS; }

int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{
    Comm cm = get_comm(c);
    std::vector<char> d = pack(b, n, t);
    raw_send(cm.ranks[size_t(dst)], cm.ctx, tag, d.data(), d.size());
    return MPI_SUCCESS;
}
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s)
{
    Comm cm = get_comm(c);
    std::vector<char> d = raw_recv(cm.ranks[size_t(src)], cm.ctx, tag);
    unpack(d, b, n, t);
    if (s) { s->MPI_SOURCE = src; s->MPI_TAG = tag; s->MPI_ERROR = MPI_SUCCESS; }
    return MPI_SUCCESS;
}
int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r)
{
    MPI_Send(b, n, t, dst, tag, c);            // buffered by the receiver's progress thread: complete on return
    Req q; q.recv = false;
    *r = new_request(q);
    return MPI_SUCCESS;
}
int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r)
{
    Comm cm = get_comm(c);
    Req q; q.recv = true; q.buf = b; q.n = n; q.t = t; q.src_world = cm.ranks[size_t(src)]; q.ctx = cm.ctx; q.tag = tag;
    *r = new_request(q);
    return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* r, MPI_Status* s)
{
    (void) s;
    if (*r == MPI_REQUEST_NULL) return MPI_SUCCESS;
    Req q;
    { std::lock_guard<std::mutex> lk(g_tab_mu); q = g_reqs[size_t(*r)]; }
    if (q.used && q.recv) {
        std::vector<char> d = raw_recv(q.src_world, q.ctx, q.tag);
        unpack(d, q.buf, q.n, q.t);
    }
    { std::lock_guard<std::mutex> lk(g_tab_mu); g_reqs[size_t(*r)].used = false; }
    *r = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void) s; for (int i = 0; i < n; ++i) MPI_Wait(&r[i], MPI_STATUS_IGNORE); return MPI_SUCCESS; }
int MPI_Request_free(MPI_Request* r)
{
    // the reference frees only send requests (already complete here); a freed receive is completed first so that its
    // message does not stay in the queue and match a later receive
    return MPI_Wait(r, MPI_STATUS_IGNORE);
}
int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype st, int dst, int stag, void* rb, int rn, MPI_Datatype rt, int src, int rtag,
                 MPI_Comm c, MPI_Status* s)
{
    MPI_Send(sb, sn, st, dst, stag, c);
    return MPI_Recv(rb, rn, rt, src, rtag, c, s);
}

int MPI_Barrier(MPI_Comm c)
{
    Comm cm = get_comm(c);
    const int tag = next_coll_tag(c);
    const int np = int(cm.ranks.size());
    char z = 0;
    if (cm.me == 0) {
        for (int r = 1; r < np; ++r) raw_recv(cm.ranks[size_t(r)], cm.ctx, tag);
        for (int r = 1; r < np; ++r) raw_send(cm.ranks[size_t(r)], cm.ctx, tag, &z, 1);
    }
    else {
        raw_send(cm.ranks[0], cm.ctx, tag, &z, 1);
        raw_recv(cm.ranks[0], cm.ctx, tag);
    }
    return MPI_SUCCESS;
}
int MPI_Bcast(void* buf, int n, MPI_Datatype t, int root, MPI_Comm c)
{
    Comm cm = get_comm(c);
    const int tag = next_coll_tag(c);
    const int np = int(cm.ranks.size());
    if (np == 1) return MPI_SUCCESS;
    if (cm.me == root) {
        std::vector<char> d = pack(buf, n, t);
        for (int r = 0; r < np; ++r) if (r != root) raw_send(cm.ranks[size_t(r)], cm.ctx, tag, d.data(), d.size());
    }
    else {
        std::vector<char> d = raw_recv(cm.ranks[size_t(root)], cm.ctx, tag);
        unpack(d, buf, n, t);
    }
    return MPI_SUCCESS;
}
static int reduce_impl(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, bool all, MPI_Comm c)
{
    Comm cm = get_comm(c);
    const int tag = next_coll_tag(c);
    const int np = int(cm.ranks.size());
    const size_t bytes = size_t(n) * type_size(t);
    const void* mine = (s == MPI_IN_PLACE) ? r : s;
    std::vector<char> acc(static_cast<const char*>(mine), static_cast<const char*>(mine) + bytes);
    if (cm.me == 0) {
        // acc = x_0 (op) x_1 (op) ... in rank order: fold from the right so that the user function's (in, inout) order holds
        std::vector<std::vector<char>> parts;
        parts.resize(size_t(np));
        parts[0] = acc;
        for (int p = 1; p < np; ++p) parts[size_t(p)] = raw_recv(cm.ranks[size_t(p)], cm.ctx, tag);
        acc = parts[size_t(np - 1)];
        for (int p = np - 2; p >= 0; --p) reduce_local(op, t, parts[size_t(p)].data(), acc.data(), n);
        if (all) { for (int p = 1; p < np; ++p) raw_send(cm.ranks[size_t(p)], cm.ctx, tag, acc.data(), bytes); }
        else if (root != 0) raw_send(cm.ranks[size_t(root)], cm.ctx, tag, acc.data(), bytes);
        if (all || root == 0) memcpy(r, acc.data(), bytes);
    }
    else {
        raw_send(cm.ranks[0], cm.ctx, tag, acc.data(), bytes);
        if (all || cm.me == root) {
            std::vector<char> d = raw_recv(cm.ranks[0], cm.ctx, tag);
            memcpy(r, d.data(), bytes);
        }
    }
    return MPI_SUCCESS;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { return reduce_impl(s, r, n, t, op, 0, true, c); }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { return reduce_impl(s, r, n, t, op, root, false, c); }
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* counts, const int* displs, MPI_Datatype rt, MPI_Comm c)
{
    Comm cm = get_comm(c);
    const int tag = next_coll_tag(c);
    const int np = int(cm.ranks.size());
    const size_t es = type_size(rt);
    const void* mine = (s == MPI_IN_PLACE) ? static_cast<char*>(r) + size_t(displs[cm.me]) * es : s;
    const size_t my_bytes = (s == MPI_IN_PLACE) ? size_t(counts[cm.me]) * es : size_t(sn) * type_size(st);
    for (int p = 0; p < np; ++p) if (p != cm.me) raw_send(cm.ranks[size_t(p)], cm.ctx, tag, mine, my_bytes);
    if (s != MPI_IN_PLACE) memcpy(static_cast<char*>(r) + size_t(displs[cm.me]) * es, mine, my_bytes);
    for (int p = 0; p < np; ++p)
        if (p != cm.me) {
            std::vector<char> d = raw_recv(cm.ranks[size_t(p)], cm.ctx, tag);
            memcpy(static_cast<char*>(r) + size_t(displs[p]) * es, d.data(), d.size());
        }
    return MPI_SUCCESS;
}

} // extern "C"
