// batch_recon_test.cpp -- mpc/tests/batchrecon_test.rs restated against the C++ BatchReconNode mirror
// (include/hbmpc_batch_recon.hpp) and the batch C ABI.  All field arithmetic runs on the GPU (no CPU fallback); the network
// is an in-process FakeNetwork (per-party FIFO inboxes of bincode-framed WrappedMessage bytes, as in
// stoffelmpc_network::fake_network).
//   test_batch_reconstruction            tests/batchrecon_test.rs:120-215  (n=4, t=1, secrets [3,6], Eval/Reveal arms)
//   test_batch_reconstruction_many       the EvalBatch/RevealBatch arms (batch_recon.rs:144-185, 332-481): n=16, t=5, 40 chunks
//   test_batch_reconstruction_byzantine  t parties send garbage in both rounds and arrive FIRST: honest parties fail on the
//                                        threshold message, retry on each later message and still open every secret
//   test_message_framing                 WrappedMessage::BatchRecon bytes round-trip, foreign variants are ignored
#include <cstdio>
#include <functional>
#include <random>

#include "hbmpc_batch_recon.hpp"

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
    std::function<std::vector<uint8_t>(size_t, const std::vector<uint8_t> &)> tamper;  // Byzantine sender: rewrites what leaves this party
    FakeNetwork(size_t id_, FakeInnerNetwork &in) : id(id_), inner(in) {}
    void send(size_t recipient, const std::vector<uint8_t> &bytes) override {
        inner.inbox[recipient].push_back(tamper ? tamper(recipient, bytes) : bytes);
    }
    void broadcast(const std::vector<uint8_t> &bytes) override {
        for (size_t j = 0; j < inner.inbox.size(); ++j) send(j, bytes);
    }
};

// generate_independent_shares(secrets, t, n): shares[i][k] = party i's share of secret k
static std::vector<std::vector<Share>> generate_independent_shares(Context &ctx, const std::vector<U256> &secrets, size_t t, size_t n, std::mt19937_64 &gen) {
    std::function<uint64_t()> rng = [&gen]() { return gen(); };
    std::vector<std::vector<Share>> by_party(n);
    for (const U256 &s : secrets) {
        std::vector<Share> sh = RobustShare::compute_shares(ctx, s, n, t, rng);
        for (size_t i = 0; i < n; ++i) by_party[i].push_back(sh[i]);
    }
    return by_party;
}

static std::vector<U256> deser_result(const std::vector<uint8_t> &bytes) {
    uint64_t len;
    REQUIRE(bytes.size() >= 8);
    std::memcpy(&len, bytes.data(), 8);
    REQUIRE(bytes.size() == 8 + 32 * len);
    std::vector<U256> v(len);
    if (len) std::memcpy(v[0].data(), bytes.data() + 8, 32 * len);
    return v;
}

// delivers messages round-robin until every honest party holds the secrets; returns the number of handler errors seen
static size_t run_network(std::vector<BatchReconNode> &nodes, std::vector<FakeNetwork> &nets, FakeInnerNetwork &inner, SessionId sid,
                          const std::vector<bool> &honest, bool reverse_order = false) {
    size_t errors = 0, idle = 0;
    const size_t n = nodes.size();
    while (idle < 2) {
        bool any = false;
        for (size_t k = 0; k < n; ++k) {
            const size_t j = reverse_order ? n - 1 - k : k;
            if (inner.inbox[j].empty()) continue;
            any = true;
            std::vector<uint8_t> raw = std::move(inner.inbox[j].front());
            inner.inbox[j].pop_front();
            if (!honest[j]) continue;  // Byzantine parties do not follow the protocol
            std::optional<BatchReconMsg> m = BatchReconMsg::decode(raw);
            if (!m) continue;          // "Malformed or unrecognized message format."
            try {
                nodes[j].process(*m, nets[j]);
            } catch (const BatchReconError &) {
                ++errors;              // "Processing failure" (the reference test logs and carries on)
            }
        }
        idle = any ? 0 : idle + 1;
    }
    for (size_t j = 0; j < n; ++j)
        if (honest[j]) REQUIRE(nodes[j].has_secrets(sid));
    return errors;
}

static void test_batch_reconstruction(Context &ctx) {
    const size_t n = 4, t = 1;
    const SessionId sid = SessionId::make(PROTOCOL_BATCH_RECON, 123, 0, 0, 111);
    std::mt19937_64 gen(7);
    const std::vector<U256> secrets = {fr_from_u64(3), fr_from_u64(6)};
    auto all_shares = generate_independent_shares(ctx, secrets, t, n, gen);
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<BatchReconNode> nodes;
    nets.reserve(n);
    nodes.reserve(n);
    for (size_t i = 0; i < n; ++i) { nets.emplace_back(i, inner); nodes.emplace_back(ctx, i, n, t, t); }
    for (size_t i = 0; i < n; ++i) nodes[i].init_batch_reconstruct(all_shares[i], sid, nets[i]);
    REQUIRE(run_network(nodes, nets, inner, sid, std::vector<bool>(n, true)) == 0);
    for (size_t i = 0; i < n; ++i) {
        std::vector<U256> got = deser_result(nodes[i].get_store(sid));
        REQUIRE(got.size() == t + 1 && got[0] == secrets[0] && got[1] == secrets[1]);
        REQUIRE(nodes[i].output.size() == 1 && nodes[i].output[0] == sid);
    }
    // error paths of the reference API
    try { nodes[0].get_store(SessionId::make(PROTOCOL_BATCH_RECON, 9, 9, 9, 9)); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::InvalidInput); }
    try { nodes[0].init_batch_reconstruct({all_shares[0][0]}, sid, nets[0]); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::InvalidInput); }
    REQUIRE(nodes[0].clear_store(sid) && nodes[0].store_len() == 0);
    std::puts("test_batch_reconstruction ok");
}

static void test_batch_reconstruction_many(Context &ctx, bool byzantine) {
    const size_t n = 16, t = 5, chunks = 40;
    const SessionId sid = SessionId::make(PROTOCOL_BATCH_RECON, 7, 1, 0, byzantine ? 2 : 1);
    std::mt19937_64 gen(byzantine ? 99 : 11);
    std::function<uint64_t()> rng = [&gen]() { return gen(); };
    std::vector<U256> secrets(chunks * (t + 1));
    for (auto &s : secrets) s = fr_rand(rng);
    auto all_shares = generate_independent_shares(ctx, secrets, t, n, gen);
    FakeInnerNetwork inner(n);
    std::vector<FakeNetwork> nets;
    std::vector<BatchReconNode> nodes;
    nets.reserve(n);
    nodes.reserve(n);
    for (size_t i = 0; i < n; ++i) { nets.emplace_back(i, inner); nodes.emplace_back(ctx, i, n, t, t); }
    std::vector<bool> honest(n, true);
    if (byzantine) {
        // parties 0..t-1 are corrupted: they flip bits in every value they send (round 1), and what they would broadcast in
        // round 2 is replaced by garbage of the right width; their messages are delivered first (lowest inbox positions)
        for (size_t i = 0; i < t; ++i) {
            honest[i] = false;
            nets[i].tamper = [](size_t recipient, const std::vector<uint8_t> &bytes) {
                std::vector<uint8_t> b = bytes;
                for (size_t off = 40 + 8; off + 32 <= b.size(); off += 32) b[off] ^= (uint8_t)(1 + recipient);  // low byte: value stays < r
                return b;
            };
        }
    }
    for (size_t i = 0; i < n; ++i) nodes[i].init_batch_reconstruct_many(all_shares[i], sid, nets[i]);
    if (byzantine) {
        for (size_t i = 0; i < t; ++i) {  // round-2 garbage from the corrupted parties, queued before any honest reveal
            std::vector<U256> junk(chunks);
            for (auto &x : junk) x = fr_rand(rng);
            BatchReconMsg m{sid, i, BatchReconMsgType::RevealBatch, detail::ser_vec(junk)};
            for (size_t j = 0; j < n; ++j) inner.inbox[j].push_back(m.encode());
        }
    }
    const size_t errors = run_network(nodes, nets, inner, sid, honest);
    if (byzantine) REQUIRE(errors > 0);  // the threshold message cannot be decoded yet: Err, then a retry per later message
    else REQUIRE(errors == 0);
    for (size_t i = 0; i < n; ++i) {
        if (!honest[i]) continue;
        std::vector<U256> got = deser_result(nodes[i].get_store(sid));
        REQUIRE(got.size() == secrets.size());
        for (size_t k = 0; k < secrets.size(); ++k) REQUIRE(got[k] == secrets[k]);
    }
    {   // late / duplicate messages of a TERMINATED session are dropped (batch_recon.rs:519-530), whatever they carry: no error, no state
        const size_t who = n - 1, before = nodes[who].store_len();
        const std::vector<uint8_t> secrets_before = nodes[who].get_store(sid);
        BatchReconMsg late1{sid, 2, BatchReconMsgType::RevealBatch, detail::ser_vec({fr_from_u64(1)})};   // wrong width: would be InvalidInput on a live session
        BatchReconMsg late2{sid, 3, BatchReconMsgType::EvalBatch, detail::ser_vec({fr_from_u64(9), fr_from_u64(9)})};
        nodes[who].process(late1, nets[who]);
        nodes[who].process(late2, nets[who]);
        REQUIRE(nodes[who].store_len() == before && nodes[who].get_store(sid) == secrets_before && nodes[who].output.size() == 1);
    }
    // reference-shaped input validation
    try { nodes[n - 1].init_batch_reconstruct_many({all_shares[n - 1][0]}, sid, nets[n - 1]); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::InvalidInput); }
    {
        BatchReconMsg bad{SessionId::make(PROTOCOL_BATCH_RECON, 8, 0, 0, 5), 3, BatchReconMsgType::EvalBatch, detail::ser_vec({})};
        try { nodes[n - 1].process(bad, nets[n - 1]); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::InvalidInput); }
        BatchReconMsg w1{bad.session_id, 3, BatchReconMsgType::EvalBatch, detail::ser_vec({fr_from_u64(1), fr_from_u64(2)})};
        BatchReconMsg w2{bad.session_id, 4, BatchReconMsgType::EvalBatch, detail::ser_vec({fr_from_u64(1)})};
        nodes[n - 1].process(w1, nets[n - 1]);
        try { nodes[n - 1].process(w2, nets[n - 1]); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::InvalidInput); }
        std::vector<uint8_t> noncanon(8 + 32, 0xff);
        const uint64_t one = 1;
        std::memcpy(noncanon.data(), &one, 8);
        BatchReconMsg w3{bad.session_id, 5, BatchReconMsgType::EvalBatch, noncanon};
        try { nodes[n - 1].process(w3, nets[n - 1]); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::ArkDeserialization); }
    }
    std::printf("test_batch_reconstruction_many%s ok (%zu handler errors on the way)\n", byzantine ? "_byzantine" : "", errors);
}

static void test_message_framing() {
    BatchReconMsg m{SessionId::make(PROTOCOL_BATCH_RECON, 123, 0, 0, 111), 3, BatchReconMsgType::RevealBatch, {1, 2, 3, 4, 5}};
    std::vector<uint8_t> raw = m.encode();
    REQUIRE(raw.size() == 4 + 16 + 8 + 4 + 8 + 5);
    REQUIRE(raw[0] == 2 && raw[1] == 0);                 // WrappedMessage variant index, u32 LE
    REQUIRE(raw[4] == 111 && raw[4 + 14] == 6);          // instance id in the low bytes, ProtocolType::BatchRecon at bits 112..120
    std::optional<BatchReconMsg> back = BatchReconMsg::decode(raw);
    REQUIRE(back && back->session_id == m.session_id && back->sender_id == 3 && back->msg_type == BatchReconMsgType::RevealBatch && back->payload == m.payload);
    raw[0] = 1;                                          // WrappedMessage::Rbc: not ours
    REQUIRE(!BatchReconMsg::decode(raw));
    raw[0] = 2;
    raw.resize(raw.size() - 1);                          // truncated payload
    REQUIRE(!BatchReconMsg::decode(raw));
    std::puts("test_message_framing ok");
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--host-only") {  // no device: wire framing and payload validation only
        test_message_framing();
        try { detail::deser_bounded_vec(std::vector<uint8_t>(4, 0), 4); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::ArkDeserialization); }
        std::vector<uint8_t> big(8 + 32, 0);
        big[0] = 200;  // claims 200 elements in a 40-byte payload
        try { detail::deser_bounded_vec(big, big.size()); REQUIRE(false); } catch (const BatchReconError &e) { REQUIRE(e.kind == BatchReconError::ArkDeserialization); }
        std::vector<U256> two = {fr_from_u64(5), fr_from_u64(6)};
        REQUIRE(detail::deser_bounded_vec(detail::ser_vec(two), 72) == two);
        std::puts("host-only checks passed");
        return 0;
    }
    Context ctx(0);
    test_message_framing();
    test_batch_reconstruction(ctx);
    test_batch_reconstruction_many(ctx, false);
    test_batch_reconstruction_many(ctx, true);
    std::puts("all batch reconstruction tests passed");
    return 0;
}
