// triple_mul_test.cpp -- the reference's triple-generation and multiplication flows restated against the C++ mirrors
// (include/hbmpc_triple_mul.hpp) over an in-process FakeNetwork; all field arithmetic runs on the GPU.
//   test_triple_gen             mpc/tests/triple_gen_test.rs shape: n parties hold random [a], [b] and double shares ([r]_t, [r]_2t); after
//                               TripleGenNode::init_batch + one batched opening of degree 2t every party holds [c] with c = a*b
//   test_mul_with_preprocessing BASELINE configs[0] / mpc/tests/node_test.rs:584-764 (mul_e2e_with_preprocessing): n = 4, t = 1,
//                               2t+1 = 3 triples from the run above, inputs 10 and 20: the products open to 100 and 400
//   test_mul_five               mpc/tests/node_test.rs:447-581 shape: five multiplications at t = 1: two full (t+1)-chunks through the
//                               batched openings, one remainder value through the reliable-broadcast path; a Byzantine party's
//                               remainder shares are corrected by the robust interpolation
#include <cstdio>
#include <functional>
#include <random>

#include "hbmpc_triple_mul.hpp"

using namespace hbmpc;

#define REQUIRE(cond)                                                            \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "%s:%d: REQUIRE(%s) failed\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                        \
        }                                                                        \
    } while (0)

struct FakeInnerNetwork {
    std::vector<std::deque<std::vector<uint8_t>>> inbox;
    std::vector<std::deque<std::pair<size_t, std::vector<uint8_t>>>> rbc_inbox;
    explicit FakeInnerNetwork(size_t n) : inbox(n), rbc_inbox(n) {}
};
struct FakeNetwork : Network, Rbc {
    size_t id;
    FakeInnerNetwork &inner;
    std::function<std::vector<uint8_t>(const std::vector<uint8_t> &)> tamper_rbc;
    FakeNetwork(size_t id_, FakeInnerNetwork &in) : id(id_), inner(in) {}
    void send(size_t recipient, const std::vector<uint8_t> &bytes) override { inner.inbox[recipient].push_back(bytes); }
    void broadcast(const std::vector<uint8_t> &bytes) override {
        for (size_t j = 0; j < inner.inbox.size(); ++j) send(j, bytes);
    }
    void init(size_t sender, SessionId, const std::vector<uint8_t> &bytes) override {   // reliable broadcast: the same bytes for everyone
        const std::vector<uint8_t> b = tamper_rbc ? tamper_rbc(bytes) : bytes;
        for (size_t j = 0; j < inner.rbc_inbox.size(); ++j) inner.rbc_inbox[j].emplace_back(sender, b);
    }
};

static std::function<uint64_t()> g_rng;

// shares[i][k] = party i's share of secret k
static std::vector<std::vector<Share>> deal(Context &ctx, const std::vector<U256> &secrets, size_t degree, size_t n) {
    std::vector<std::vector<Share>> by_party(n);
    for (const U256 &s : secrets) {
        std::vector<Share> sh = RobustShare::compute_shares(ctx, s, n, degree, g_rng);
        for (size_t i = 0; i < n; ++i) by_party[i].push_back(sh[i]);
    }
    return by_party;
}
static U256 open(Context &ctx, const std::vector<Share> &shares, size_t n, size_t t) { return RobustShare::recover_secret(ctx, shares, n, t).second; }
static U256 fr_mul(Context &ctx, const U256 &a, const U256 &b) {
    U256 out;
    check(hbmpc_elementwise(ctx.get(), 2, 1, a.data(), b.data(), out.data()));
    return out;
}

template <class Node>
static void pump(std::vector<Node> &nodes, std::vector<FakeNetwork> &nets, FakeInnerNetwork &inner, const std::function<void(size_t)> &rbc_hook = nullptr) {
    for (size_t idle = 0; idle < 2;) {
        bool any = false;
        for (size_t j = 0; j < nodes.size(); ++j) {
            if (!inner.inbox[j].empty()) {
                any = true;
                std::vector<uint8_t> raw = std::move(inner.inbox[j].front());
                inner.inbox[j].pop_front();
                std::optional<BatchReconMsg> m = BatchReconMsg::decode(raw);
                if (m) {
                    try { nodes[j].process(*m, nets[j]); } catch (const BatchReconError &) { /* retried on the next arrival */ }
                }
            }
            if (rbc_hook && !inner.rbc_inbox[j].empty()) { any = true; rbc_hook(j); }
        }
        idle = any ? 0 : idle + 1;
    }
}

static std::vector<std::vector<ShamirBeaverTriple>> test_triple_gen(Context &ctx, size_t n, size_t t, size_t groups) {
    const size_t count = groups * (2 * t + 1);
    std::vector<U256> a(count), b(count), r(count);
    for (size_t i = 0; i < count; ++i) { a[i] = fr_rand(g_rng); b[i] = fr_rand(g_rng); r[i] = fr_rand(g_rng); }
    auto sa = deal(ctx, a, t, n), sb = deal(ctx, b, t, n), rt = deal(ctx, r, t, n), r2t = deal(ctx, r, 2 * t, n);
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<TripleGenNode> nodes;
    nets.reserve(n);
    nodes.reserve(n);
    for (size_t i = 0; i < n; ++i) { nets.emplace_back(i, inner); nodes.emplace_back(ctx, i, n, t); }
    const SessionId sid = SessionId::make(PROTOCOL_TRIPLE, 5, 0, 0, 111);
    for (size_t i = 0; i < n; ++i) {
        std::vector<DoubleShamirShare> pairs(count);
        for (size_t k = 0; k < count; ++k) pairs[k] = DoubleShamirShare{rt[i][k], r2t[i][k]};
        nodes[i].init_batch(sa[i], sb[i], pairs, sid, nets[i]);
    }
    pump(nodes, nets, inner);
    std::vector<std::vector<ShamirBeaverTriple>> out(n);
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(nodes[i].output.count(sid) == 1);
        out[i] = nodes[i].output[sid];
        REQUIRE(out[i].size() == count);
    }
    for (size_t k = 0; k < count; ++k) {   // open every triple: c == a*b, shares of degree t with the party's id
        std::vector<Share> ca(n), cb(n), cc(n);
        for (size_t i = 0; i < n; ++i) { ca[i] = out[i][k].a; cb[i] = out[i][k].b; cc[i] = out[i][k].mult; REQUIRE(cc[i].degree == t && cc[i].id == i); }
        REQUIRE(open(ctx, ca, n, t) == a[k] && open(ctx, cb, n, t) == b[k]);
        REQUIRE(open(ctx, cc, n, t) == fr_mul(ctx, a[k], b[k]));
    }
    // reference-shaped input validation: a group that is not a multiple of 2t+1
    try {
        std::vector<DoubleShamirShare> pairs(1, DoubleShamirShare{rt[0][0], r2t[0][0]});
        nodes[0].init_batch({sa[0][0]}, {sb[0][0]}, pairs, SessionId::make(PROTOCOL_TRIPLE, 6, 0, 0, 111), nets[0]);
        REQUIRE(2 * t + 1 == 1);
    } catch (const TripleGenError &e) { REQUIRE(e.kind == TripleGenError::NotEnoughPreprocessing); }
    std::printf("test_triple_gen n=%zu t=%zu: %zu triples ok\n", n, t, count);
    return out;
}

static void test_mul(Context &ctx, size_t n, size_t t, const std::vector<uint64_t> &xs, const std::vector<uint64_t> &ys,
                     const std::vector<std::vector<ShamirBeaverTriple>> &triples, int byzantine) {
    const size_t m = xs.size();
    std::vector<U256> xv(m), yv(m);
    for (size_t i = 0; i < m; ++i) { xv[i] = fr_from_u64(xs[i]); yv[i] = fr_from_u64(ys[i]); }
    auto sx = deal(ctx, xv, t, n), sy = deal(ctx, yv, t, n);
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<Multiply> nodes;
    nets.reserve(n);
    nodes.reserve(n);
    for (size_t i = 0; i < n; ++i) { nets.emplace_back(i, inner); nodes.emplace_back(ctx, i, n, t); }
    if (byzantine >= 0)   // this party's remainder shares are garbage (same garbage for everyone: it went through reliable broadcast)
        nets[byzantine].tamper_rbc = [](const std::vector<uint8_t> &b) { std::vector<uint8_t> c = b; for (size_t o = 16; o < c.size(); o += 32) c[o] ^= 0x3c; return c; };
    const SessionId sid = SessionId::make(PROTOCOL_MUL, 9, 0, 0, 111);
    for (size_t i = 0; i < n; ++i) {
        std::vector<ShamirBeaverTriple> tr(triples[i].begin(), triples[i].begin() + m);
        nodes[i].init(sid, sx[i], sy[i], tr, nets[i], nets[i]);
    }
    pump(nodes, nets, inner, [&](size_t j) {
        auto msg = std::move(inner.rbc_inbox[j].front());
        inner.rbc_inbox[j].pop_front();
        nodes[j].rbc_deliver(msg.first, msg.second);
    });
    for (size_t k = 0; k < m; ++k) {
        std::vector<Share> z(n);
        for (size_t i = 0; i < n; ++i) {
            REQUIRE(nodes[i].output.count(sid) == 1 && nodes[i].output[sid].size() == m);
            z[i] = nodes[i].output[sid][k];
            REQUIRE(z[i].degree == t && z[i].id == i);       // node_test.rs:552-556
        }
        REQUIRE(open(ctx, z, n, t) == fr_from_u64(xs[k] * ys[k]));   // node_test.rs:575-580, :757-760
    }
    try { nodes[0].init(sid, sx[0], {sy[0][0]}, {}, nets[0], nets[0]); REQUIRE(m == 1); } catch (const MulError &e) { REQUIRE(e.kind == MulError::InvalidInput); }
    std::printf("test_mul n=%zu t=%zu: %zu multiplications ok%s\n", n, t, m, byzantine >= 0 ? " (one Byzantine remainder message)" : "");
}

int main() {
    Context ctx(0);
    std::mt19937_64 gen(2024);
    g_rng = [&gen]() { return gen(); };
    // BASELINE configs[0]: 4 parties, t = 1
    auto triples = test_triple_gen(ctx, 4, 1, 2);                       // 2 groups of 2t+1 = 3 triples
    test_mul(ctx, 4, 1, {10, 20}, {10, 20}, triples, -1);               // 100, 400 (mul_e2e_with_preprocessing)
    test_mul(ctx, 4, 1, {1, 2, 3, 4, 5}, {6, 7, 8, 9, 10}, triples, -1);  // five multiplications: two batched chunks + one remainder value
    test_mul(ctx, 4, 1, {1, 2, 3, 4, 5}, {6, 7, 8, 9, 10}, triples, 2);
    auto big = test_triple_gen(ctx, 16, 5, 4);                          // 44 triples at n = 16, t = 5
    test_mul(ctx, 16, 5, {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47}, {2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28}, big, 3);
    std::puts("all triple generation / multiplication tests passed");
    return 0;
}
