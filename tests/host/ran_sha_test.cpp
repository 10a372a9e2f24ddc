// ran_sha_test.cpp -- the reference's RanSha flow (share_gen.rs, mpc/tests/share_gen_test.rs) restated against the C++ RanShaNode mirror
// (include/hbmpc_ran_sha.hpp) and the batch C ABI; all field arithmetic on the GPU, in-process FakeNetwork, the reliable broadcast
// of the verdicts replaced by direct delivery.
//   test_message_framing         WrappedMessage::RanSha bytes round-trip (host only)
//   run(n, t, B, honest)         every party deals B random secrets (K1, one call), receives n share vectors, applies the n x n
//                                hyperinvertible matrix (K2, one call), the first 2t parties robustly recover r_i for every batch
//                                column (K3/K4, one call) and test its degree; everybody outputs B*(n-2t) shares that form
//                                degree-t sharings of common random values
//   run(..., bad dealer)         a dealer whose polynomial has degree t+1: the r_i are not degree-t sharings, every verifier says false
#include <cstdio>
#include <random>

#include "hbmpc_ran_sha.hpp"

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
    RanShaMessage m;
    m.sender_id = 7;
    m.msg_type = RanShaMessage::ReconstructMessage;
    m.session_id = SessionId::make(PROTOCOL_RANSHA, 11, 0, 2, 3);
    m.kind = RanShaMessage::ReconstructSharesBatch;
    m.bytes = {1, 0, 0, 0, 0, 0, 0, 0};
    m.bytes.resize(8 + 48, 0);
    std::vector<uint8_t> raw = m.encode();
    REQUIRE(raw.size() == 4 + 8 + 4 + 16 + 4 + 8 + 56 && raw[0] == 4);
    auto back = RanShaMessage::decode(raw);
    REQUIRE(back && back->sender_id == 7 && back->msg_type == RanShaMessage::ReconstructMessage && back->session_id == m.session_id &&
            back->kind == RanShaMessage::ReconstructSharesBatch && back->bytes == m.bytes);
    RanShaMessage o;
    o.sender_id = 1; o.msg_type = RanShaMessage::OutputMessage; o.session_id = m.session_id; o.kind = RanShaMessage::Output; o.ok = true;
    auto ob = RanShaMessage::decode(o.encode());
    REQUIRE(ob && ob->kind == RanShaMessage::Output && ob->ok && ob->msg_type == RanShaMessage::OutputMessage);
    raw[0] = 0;  // WrappedMessage::RanDouSha: not ours
    REQUIRE(!RanShaMessage::decode(raw));
    raw[0] = 4;
    raw.push_back(0);
    REQUIRE(!RanShaMessage::decode(raw));
    std::puts("test_message_framing ok");
}

static void run(Context &ctx, size_t n, size_t t, size_t B, int bad_dealer) {
    std::mt19937_64 gen(n * 977 + t * 13 + B + (bad_dealer >= 0 ? 5 : 0));
    std::function<uint64_t()> rng = [&gen]() { return gen(); };
    const SessionId sid = SessionId::make(PROTOCOL_RANSHA, 9, 0, 0, (uint32_t)(n * 10 + B));
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<RanShaNode> nodes;
    nets.reserve(n);
    nodes.reserve(n);
    std::vector<RanShaMessage> verdicts;
    for (size_t i = 0; i < n; ++i) {
        nets.emplace_back(i, inner);
        nodes.emplace_back(ctx, i, n, t);
        nodes.back().broadcast_output = [&verdicts](const RanShaMessage &m) { verdicts.push_back(m); };
    }
    for (size_t i = 0; i < n; ++i) {
        if ((int)i != bad_dealer) {
            nodes[i].init_batch(sid, B, rng, nets[i]);
            continue;
        }
        // Byzantine dealer: polynomials of degree t+1 labelled as degree t
        std::vector<std::vector<Share>> per_recipient(n);
        for (size_t b = 0; b < B; ++b) {
            std::vector<Share> sh = NonRobustShare::compute_shares(ctx, fr_rand(rng), n, t + 1, rng);
            for (size_t j = 0; j < n; ++j) per_recipient[j].push_back(Share{sh[j].share, j, t});
        }
        for (size_t j = 0; j < n; ++j) {
            RanShaMessage m;
            m.sender_id = i; m.msg_type = RanShaMessage::ShareMessage; m.session_id = sid;
            m.kind = B == 1 ? RanShaMessage::Share : RanShaMessage::SharesBatch;
            std::vector<uint8_t> bytes(B == 1 ? 48 : 8 + 48 * B);
            size_t off = 0;
            if (B > 1) { const uint64_t len = B; std::memcpy(bytes.data(), &len, 8); off = 8; }
            for (size_t b = 0; b < B; ++b) ReconstructionMessage::put_share(bytes.data() + off + 48 * b, per_recipient[j][b]);
            m.bytes = bytes;
            nets[i].send(j, m.encode());
        }
    }
    // deliver until quiet: share messages trigger the matrix apply, whose reconstruct messages reach the verifiers 0..2t-1
    size_t idle = 0, checks = 0;
    while (idle < 2) {
        bool any = false;
        for (size_t j = 0; j < n; ++j) {
            if (inner.inbox[j].empty()) continue;
            any = true;
            auto m = RanShaMessage::decode(inner.inbox[j].front());
            inner.inbox[j].pop_front();
            REQUIRE(m.has_value());
            if (m->msg_type == RanShaMessage::ReconstructMessage) {
                REQUIRE(j < 2 * t);
                std::optional<bool> v = nodes[j].reconstruction_handler(*m);
                if (v) {
                    ++checks;
                    REQUIRE(*v == (bad_dealer < 0));
                }
            } else {
                nodes[j].process(*m, nets[j]);
            }
        }
        idle = any ? 0 : idle + 1;
    }
    // every verifier checks on the (2t+1)-th reconstruct message and on each later one
    REQUIRE(checks == 2 * t * (n - 2 * t) && verdicts.size() == checks);
    if (bad_dealer >= 0) {
        for (const RanShaMessage &v : verdicts) REQUIRE(!v.ok);
        try { nodes[n - 1].output_handler(verdicts[0]); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::Abort); }
        REQUIRE(nodes[n - 1].get_or_create_store(sid).state != RanShaStore::Finished);
        std::printf("run n=%zu t=%zu B=%zu bad dealer ok\n", n, t, B);
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        for (const RanShaMessage &v : verdicts) nodes[i].output_handler(v);
        const RanShaStore &st = nodes[i].get_or_create_store(sid);
        REQUIRE(st.state == RanShaStore::Finished && st.protocol_output.size() == B * (n - 2 * t));
    }
    for (size_t k = 0; k < B * (n - 2 * t); ++k) {
        std::vector<Share> sh(n);
        for (size_t i = 0; i < n; ++i) sh[i] = Share{nodes[i].get_or_create_store(sid).protocol_output[k].share, i, t};
        auto r = RobustShare::recover_secret(ctx, sh, n, t);
        REQUIRE(r.first.size() == t + 1);
        std::vector<Share> sub(sh.begin() + 1, sh.begin() + 1 + 2 * t + 1);  // any 2t+1 of them open the same value
        REQUIRE(RobustShare::recover_secret(ctx, sub, n, t).second == r.second);
    }
    // reference-shaped error paths
    {
        RanShaMessage w;
        w.sender_id = 1; w.msg_type = RanShaMessage::ShareMessage; w.kind = RanShaMessage::Share;
        w.session_id = SessionId::make(PROTOCOL_RANSHA, 9, 3, 0, 1);  // sub_id != 0
        w.bytes.assign(48, 0);
        try { nodes[0].receive_shares_handler(w, nets[0]); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::SessionIdError); }
        try { nodes[0].reconstruction_handler(w); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::SessionIdError); }
        w.session_id = SessionId::make(PROTOCOL_RANSHA, 10, 0, 0, 1);
        w.sender_id = n;
        try { nodes[0].receive_shares_handler(w, nets[0]); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::InvalidPartyId); }
        w.sender_id = 1;
        ReconstructionMessage::put_share(w.bytes.data(), Share{fr_from_u64(3), 2, t});  // addressed to party 2, received by party 0
        try { nodes[0].receive_shares_handler(w, nets[0]); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::ShareErr && e.code == HBMPC_ID_MISMATCH); }
        RanShaMessage o;
        o.sender_id = 2 * t; o.msg_type = RanShaMessage::OutputMessage; o.session_id = sid; o.kind = RanShaMessage::Output; o.ok = true;
        try { nodes[0].output_handler(o); REQUIRE(false); } catch (const RanShaError &e) { REQUIRE(e.kind == RanShaError::InvalidPartyId); }
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
    run(ctx, 4, 1, 1, -1);
    run(ctx, 4, 1, 3, -1);
    run(ctx, 16, 5, 4, -1);
    run(ctx, 64, 21, 2, -1);
    run(ctx, 4, 1, 2, 0);
    run(ctx, 16, 5, 3, 7);
    std::puts("all RanSha tests passed");
    return 0;
}
