// double_share_test.cpp -- DoubleShareNode -> RanDouShaNode chained through the C++ mirrors (include/hbmpc_double_share.hpp,
// include/hbmpc_ran_dou_sha.hpp): every party deals B double sharings (two batched K1 calls), the collected double shares
// ([batch][dealer]) are exactly RanDouSha's input, whose checkers then accept; the outputs open consistently.  GPU arithmetic only.
#include <cstdio>
#include <random>

#include "hbmpc_double_share.hpp"

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
    explicit FakeInnerNetwork(size_t n) : inbox(n) {}
};
struct FakeNetwork : Network {
    size_t id;
    FakeInnerNetwork &inner;
    FakeNetwork(size_t id_, FakeInnerNetwork &in) : id(id_), inner(in) {}
    void send(size_t recipient, const std::vector<uint8_t> &bytes) override { inner.inbox[recipient].push_back(bytes); }
    void broadcast(const std::vector<uint8_t> &bytes) override {
        for (auto &q : inner.inbox) q.push_back(bytes);
    }
};

static void test_message_framing() {
    DouShaMessage m;
    m.sender_id = 2;
    m.session_id = SessionId::make(PROTOCOL_DOUSHA, 3, 0, 0, 4);
    m.kind = DouShaMessage::Shares;
    m.bytes.assign(8 + 96, 7);
    std::vector<uint8_t> raw = m.encode();
    REQUIRE(raw.size() == 4 + 8 + 16 + 4 + 8 + 104 && raw[0] == 5);
    auto back = DouShaMessage::decode(raw);
    REQUIRE(back && back->sender_id == 2 && back->session_id == m.session_id && back->kind == DouShaMessage::Shares && back->bytes == m.bytes);
    raw[0] = 4;
    REQUIRE(!DouShaMessage::decode(raw));
    std::puts("test_message_framing ok");
}

static void run(Context &ctx, size_t n, size_t t, size_t B) {
    std::mt19937_64 gen(n * 31 + B);
    std::function<uint64_t()> rng = [&gen]() { return gen(); };
    const SessionId sid = SessionId::make(PROTOCOL_DOUSHA, 1, 0, 0, (uint32_t)(n + B));
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<DoubleShareNode> dealers;
    nets.reserve(n);
    dealers.reserve(n);
    for (size_t i = 0; i < n; ++i) { nets.emplace_back(i, inner); dealers.emplace_back(ctx, i, n, t); }
    for (size_t i = 0; i < n; ++i) dealers[i].init_batch(sid, B, rng, nets[i]);
    for (size_t j = 0; j < n; ++j) {
        REQUIRE(inner.inbox[j].size() == n);
        bool done = false;
        while (!inner.inbox[j].empty()) {
            auto m = DouShaMessage::decode(inner.inbox[j].front());
            inner.inbox[j].pop_front();
            REQUIRE(m.has_value());
            done = dealers[j].receive_double_shares_handler(*m);
        }
        REQUIRE(done && dealers[j].get_or_create_store(sid).protocol_output.size() == B * n);
    }
    // each dealt pair opens to one value with degrees t and 2t
    for (size_t b = 0; b < B; ++b)
        for (size_t dealer = 0; dealer < n; dealer += (n > 8 ? 5 : 1)) {
            std::vector<Share> st(n), s2(n);
            for (size_t i = 0; i < n; ++i) {
                const DoubleShamirShare &d = dealers[i].get_or_create_store(sid).protocol_output[b * n + dealer];
                st[i] = d.degree_t;
                s2[i] = d.degree_2t;
                REQUIRE(st[i].id == i && s2[i].id == i);
            }
            auto rt = NonRobustShare::recover_secret(ctx, st, n), r2 = NonRobustShare::recover_secret(ctx, s2, n);
            REQUIRE(rt.second == r2.second && rt.first.size() == t + 1 && r2.first.size() == 2 * t + 1);
        }
    // the collected double shares are RanDouSha's input
    const SessionId rsid = SessionId::make(PROTOCOL_RANDOUSHA, 2, 0, 0, (uint32_t)(n + B));
    std::vector<RanDouShaNode> nodes;
    nodes.reserve(n);
    std::vector<RanDouShaMessage> verdicts;
    for (size_t i = 0; i < n; ++i) {
        nodes.emplace_back(ctx, i, n, t);
        nodes.back().broadcast_output = [&verdicts](const RanDouShaMessage &m) { verdicts.push_back(m); };
        std::vector<std::vector<Share>> in_t(B, std::vector<Share>(n)), in_2t(B, std::vector<Share>(n));
        const auto &out = dealers[i].get_or_create_store(sid).protocol_output;
        for (size_t b = 0; b < B; ++b)
            for (size_t dealer = 0; dealer < n; ++dealer) { in_t[b][dealer] = out[b * n + dealer].degree_t; in_2t[b][dealer] = out[b * n + dealer].degree_2t; }
        nodes[i].init_batch(in_t, in_2t, rsid, nets[i]);
    }
    for (size_t j = t + 1; j < n; ++j)
        while (!inner.inbox[j].empty()) {
            auto m = RanDouShaMessage::decode(inner.inbox[j].front());
            inner.inbox[j].pop_front();
            REQUIRE(m.has_value());
            std::optional<bool> v = nodes[j].reconstruction_handler(*m);
            if (v) REQUIRE(*v);
        }
    REQUIRE(verdicts.size() == n - (t + 1));
    for (size_t i = 0; i < n; ++i) {
        for (const auto &v : verdicts) nodes[i].output_handler(v);
        REQUIRE(nodes[i].store(rsid).finished && nodes[i].store(rsid).protocol_output.size() == B * (t + 1));
    }
    // reference-shaped error path: a double share addressed to somebody else
    {
        DouShaMessage w;
        w.sender_id = 1;
        w.session_id = SessionId::make(PROTOCOL_DOUSHA, 77, 0, 0, 1);
        w.kind = DouShaMessage::Share;
        w.bytes.assign(96, 0);
        ReconstructionMessage::put_share(w.bytes.data(), Share{fr_from_u64(1), 3, 2 * t});
        ReconstructionMessage::put_share(w.bytes.data() + 48, Share{fr_from_u64(1), 3, t});
        try { dealers[0].receive_double_shares_handler(w); REQUIRE(false); } catch (const DouShaError &e) { REQUIRE(e.kind == DouShaError::ShareErr && e.code == HBMPC_ID_MISMATCH); }
    }
    std::printf("run n=%zu t=%zu B=%zu ok\n", n, t, B);
}

int main(int argc, char **argv) {
    test_message_framing();
    if (argc > 1 && std::string(argv[1]) == "--host-only") {
        std::puts("host-only checks passed");
        return 0;
    }
    Context ctx(0);
    run(ctx, 4, 1, 1);
    run(ctx, 4, 1, 3);
    run(ctx, 16, 5, 2);
    run(ctx, 64, 21, 2);
    std::puts("all DoubleShare tests passed");
    return 0;
}
