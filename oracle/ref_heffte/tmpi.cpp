// tmpi.cpp -- TEST INFRASTRUCTURE ONLY (oracle/): thread-backed implementation of the MPI subset in mpi.h.
// Ranks are threads of one process.  Point-to-point messages are matched against posted receives (the sender
// copies straight into the receiver's buffer) or parked in an unexpected-message queue (eager copy), so the
// blocking MPI_Send of heFFTe's p2p reshape cannot deadlock; collectives exchange buffer pointers through
// per-communicator slots between two barriers and every rank copies its own incoming blocks.
#include "mpi.h"

#include <pthread.h>
#include <sched.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct tmpi_req {
    bool done = false;
    bool is_recv = false;
    void* buf = nullptr;
    size_t bytes = 0;
    int src = -1, tag = 0;    // ranks are communicator ranks
    struct Endpoint* ep = nullptr;
};

struct Unexpected {
    int src, tag;
    std::vector<char> data;
};

struct Endpoint {   // one per (communicator, member)
    std::mutex mu;
    std::condition_variable cv;
    std::list<tmpi_req*> posted;
    std::deque<Unexpected> unexpected;
};

struct tmpi_group {
    std::vector<int> world;   // world ranks of the members, in group order
};

struct tmpi_comm {
    std::vector<int> world;   // world rank of member i
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    long long gen = 0;
    int refs = 0;
    std::vector<const void*> s_buf;
    std::vector<const int*> s_cnt, s_dsp;
    std::vector<const void*> s_aux;
    std::vector<std::unique_ptr<Endpoint>> ep;
    explicit tmpi_comm(std::vector<int> w) : world(std::move(w))
    {
        const size_t n = world.size();
        s_buf.assign(n, nullptr); s_cnt.assign(n, nullptr); s_dsp.assign(n, nullptr); s_aux.assign(n, nullptr);
        for (size_t i = 0; i < n; i++) ep.emplace_back(new Endpoint);
        refs = (int)n;
    }
    int size() const { return (int)world.size(); }
    void barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        const long long g = gen;
        if (++arrived == size()) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

static tmpi_comm* g_world = nullptr;
static std::vector<int> g_cpus;   // optional: rank i is pinned to CPU g_cpus[i % size] (one entry per physical core)
static thread_local int tl_rank = 0;
static std::mutex g_job_mu;

static int rank_in(tmpi_comm* c)
{
    for (int i = 0; i < c->size(); i++)
        if (c->world[i] == tl_rank) return i;
    return MPI_UNDEFINED;
}
static inline size_t tsize(MPI_Datatype t) { return (size_t)(t & 0xff); }

extern "C" {

MPI_Comm tmpi_world(void) { return g_world; }
int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
double MPI_Wtime(void) { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int MPI_Barrier(MPI_Comm c) { c->barrier(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = rank_in(c); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = c->size(); return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { *g = new tmpi_group{c->world}; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group g, int n, const int ranks[], MPI_Group* out)
{
    tmpi_group* r = new tmpi_group;
    for (int i = 0; i < n; i++) r->world.push_back(g->world[ranks[i]]);
    *out = r;
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group* g) { delete *g; *g = nullptr; return MPI_SUCCESS; }

// collective over `c`: members of `g` get the new communicator (one shared object per distinct group), the others MPI_COMM_NULL
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* out)
{
    const int me = rank_in(c);
    bool member = false;
    for (int w : g->world) member = member || w == tl_rank;
    c->s_aux[me] = g;
    c->barrier();
    tmpi_comm* made = nullptr;
    if (member && g->world[0] == tl_rank) made = new tmpi_comm(g->world);   // the first member of the group allocates it
    c->s_buf[me] = made;
    c->barrier();
    *out = MPI_COMM_NULL;
    if (member) {
        for (int i = 0; i < c->size(); i++)
            if (c->world[i] == g->world[0]) *out = (MPI_Comm)c->s_buf[i];
    }
    c->barrier();
    return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* c)
{
    if (!c || !*c || *c == g_world) return MPI_SUCCESS;
    bool last;
    {
        std::lock_guard<std::mutex> lk((*c)->mu);
        last = --(*c)->refs == 0;
    }
    if (last) delete *c;
    *c = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Allgather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, MPI_Comm c)
{
    const int me = rank_in(c);
    c->s_buf[me] = sb;
    c->barrier();
    const size_t nb = (size_t)rc * tsize(rt);
    (void)sc; (void)st;
    for (int i = 0; i < c->size(); i++) memcpy((char*)rb + (size_t)i * nb, c->s_buf[i], nb);
    c->barrier();
    return MPI_SUCCESS;
}
int MPI_Alltoall(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, MPI_Comm c)
{
    const int me = rank_in(c);
    c->s_buf[me] = sb;
    c->barrier();
    const size_t nb = (size_t)rc * tsize(rt), sbk = (size_t)sc * tsize(st);
    for (int i = 0; i < c->size(); i++) memcpy((char*)rb + (size_t)i * nb, (const char*)c->s_buf[i] + (size_t)me * sbk, nb);
    c->barrier();
    return MPI_SUCCESS;
}
int MPI_Alltoallv(const void* sb, const int scnt[], const int sdsp[], MPI_Datatype st, void* rb, const int rcnt[], const int rdsp[],
                  MPI_Datatype rt, MPI_Comm c)
{
    const int me = rank_in(c);
    c->s_buf[me] = sb; c->s_cnt[me] = scnt; c->s_dsp[me] = sdsp;
    c->barrier();
    for (int i = 0; i < c->size(); i++) {
        const size_t nb = (size_t)rcnt[i] * tsize(rt);
        if (nb) memcpy((char*)rb + (size_t)rdsp[i] * tsize(rt), (const char*)c->s_buf[i] + (size_t)c->s_dsp[i][me] * tsize(st), nb);
    }
    c->barrier();
    return MPI_SUCCESS;
}

static void deliver(tmpi_comm* c, const void* buf, size_t bytes, int dest, int tag)
{
    const int me = rank_in(c);
    Endpoint& ep = *c->ep[dest];
    std::unique_lock<std::mutex> lk(ep.mu);
    for (auto it = ep.posted.begin(); it != ep.posted.end(); ++it) {
        tmpi_req* r = *it;
        if ((r->src == me || r->src == MPI_ANY_SOURCE) && (r->tag == tag || r->tag == MPI_ANY_TAG)) {
            ep.posted.erase(it);
            lk.unlock();
            memcpy(r->buf, buf, bytes < r->bytes ? bytes : r->bytes);   // the receive buffer is private to this request
            lk.lock();
            r->done = true;
            ep.cv.notify_all();
            return;
        }
    }
    Unexpected u{me, tag, std::vector<char>((const char*)buf, (const char*)buf + bytes)};
    ep.unexpected.push_back(std::move(u));
    ep.cv.notify_all();
}
int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{
    deliver(c, buf, (size_t)count * tsize(t), dest, tag);
    return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req)
{
    deliver(c, buf, (size_t)count * tsize(t), dest, tag);
    tmpi_req* r = new tmpi_req;
    r->done = true;
    *req = r;
    return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Request* req)
{
    const int me = rank_in(c);
    Endpoint& ep = *c->ep[me];
    tmpi_req* r = new tmpi_req;
    r->is_recv = true; r->buf = buf; r->bytes = (size_t)count * tsize(t); r->src = source; r->tag = tag; r->ep = &ep;
    std::lock_guard<std::mutex> lk(ep.mu);
    for (auto it = ep.unexpected.begin(); it != ep.unexpected.end(); ++it) {
        if ((source == it->src || source == MPI_ANY_SOURCE) && (tag == it->tag || tag == MPI_ANY_TAG)) {
            memcpy(buf, it->data.data(), it->data.size() < r->bytes ? it->data.size() : r->bytes);
            ep.unexpected.erase(it);
            r->done = true;
            *req = r;
            return MPI_SUCCESS;
        }
    }
    ep.posted.push_back(r);
    *req = r;
    return MPI_SUCCESS;
}
// Pending requests of one call are receives posted on this rank's endpoint of one communicator (sends complete at
// once), so completion is checked and awaited under that endpoint's lock: no wake-up can be missed.
int MPI_Waitany(int count, MPI_Request reqs[], int* index, MPI_Status*)
{
    Endpoint* ep = nullptr;
    bool any = false;
    for (int i = 0; i < count; i++) {
        if (!reqs[i]) continue;
        any = true;
        if (reqs[i]->is_recv) ep = reqs[i]->ep;
    }
    if (!any) { *index = MPI_UNDEFINED; return MPI_SUCCESS; }
    auto take = [&](int i) { delete reqs[i]; reqs[i] = MPI_REQUEST_NULL; *index = i; return MPI_SUCCESS; };
    for (int i = 0; i < count; i++)
        if (reqs[i] && !reqs[i]->is_recv) return take(i);
    std::unique_lock<std::mutex> lk(ep->mu);
    for (;;) {
        for (int i = 0; i < count; i++)
            if (reqs[i] && reqs[i]->done) { lk.unlock(); return take(i); }
        ep->cv.wait(lk);
    }
}
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status*)
{
    for (int i = 0; i < count; i++) {
        if (!reqs[i]) continue;
        if (reqs[i]->is_recv) {
            std::unique_lock<std::mutex> lk(reqs[i]->ep->mu);
            reqs[i]->ep->cv.wait(lk, [&] { return reqs[i]->done; });
        }
        delete reqs[i];
        reqs[i] = MPI_REQUEST_NULL;
    }
    return MPI_SUCCESS;
}

int tmpi_set_cpus(const int* cpus, int n)
{
    std::lock_guard<std::mutex> job(g_job_mu);
    g_cpus.assign(cpus, cpus + (n > 0 ? n : 0));
    return 0;
}

int tmpi_run(int nranks, void (*fn)(void*), void* arg)
{
    if (nranks < 1) return -1;
    std::lock_guard<std::mutex> job(g_job_mu);   // one job at a time: MPI_COMM_WORLD is process-global
    std::vector<int> w(nranks);
    for (int i = 0; i < nranks; i++) w[i] = i;
    g_world = new tmpi_comm(w);
    std::vector<std::thread> th;
    for (int i = 0; i < nranks; i++)
        th.emplace_back([=] {
            tl_rank = i;
            if (!g_cpus.empty()) {   // the rank allocates and first-touches its own slab after this: NUMA-local
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(g_cpus[i % g_cpus.size()], &set);
                pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
            }
            fn(arg);
        });
    for (auto& t : th) t.join();
    delete g_world;
    g_world = nullptr;
    return 0;
}
}
