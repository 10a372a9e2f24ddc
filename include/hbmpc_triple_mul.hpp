// hbmpc_triple_mul.hpp -- C++17 host-side mirrors of the reference's TripleGenNode and Multiply over the batch C ABI (SURVEY.md 8f
// N2 / N3): the callers of the hot path in Beaver-triple preprocessing and in online multiplication.
//
//   mpc/src/honeybadger/triple_gen/triple_generation.rs   init_batch :304-362 (a*b - r_2t per triple, one batched opening of degree 2t),
//                                                          try_finalize_triple_gen :165-232 (c = r_t + opened), BatchRecon of degree 2t :61
//   mpc/src/honeybadger/mul/multiplication.rs              init :342-484 (a-x, b-y; full (t+1)-chunks through two batched openings, the
//                                                          remaining < t+1 values through a reliably broadcast ReconstructionMessage),
//                                                          reconstruct_rbc :101-140, finalize_mul :57-99
//   mpc/src/honeybadger/triple_gen/mod.rs, mul/mod.rs      ShamirBeaverTriple, error variants
// Names, thresholds and error behaviour follow the reference; what changes is where the share algebra runs: every elementwise
// operation on share vectors (share_mul, Sub, Add, Mul<F>: common/mod.rs:167-300) runs on the device (K5) -- the three multi-operator
// steps (a*b - r_2t; a-x and b-y; the Beaver product share) as ONE hbmpc_share_algebra_fused pass each, the rest as hbmpc_elementwise --, the
// openings are the BatchReconNode mirror (one device call per message), the reconstruction of the remainder values one
// hbmpc_robust_interpolate_batch call.  No field arithmetic happens in this header; there is no CPU fallback.  Reliable broadcast
// (Avid / Bracha) is host control flow and out of scope: `Rbc` below is the interface the mirror needs from it (deliver the same bytes
// to every party), tests plug the FakeNetwork in.
#pragma once
#include <set>

#include "hbmpc_batch_recon.hpp"

namespace hbmpc {

struct DoubleShamirShare { Share degree_t, degree_2t; };          // double_share/mod.rs
struct ShamirBeaverTriple { Share a, b, mult; };                  // triple_gen/mod.rs
inline constexpr uint8_t PROTOCOL_TRIPLE = 4, PROTOCOL_MUL = 7;   // ProtocolType::{Triple, Mul} (honeybadger/mod.rs)

struct TripleGenError : std::runtime_error {
    enum Kind { NotEnoughPreprocessing, ShareErr, BatchRecon, NoSuchSessionId } kind;
    TripleGenError(Kind k, const std::string &what) : std::runtime_error(what), kind(k) {}
};
struct MulError : std::runtime_error {
    enum Kind { InvalidInput, ShareErr, BatchRecon, Interpolate } kind;
    MulError(Kind k, const std::string &what) : std::runtime_error(what), kind(k) {}
};

namespace detail {
// out[i] = a[i] (op) b[i] on the device: op 0 add, 1 sub, 2 mul (K5)
inline std::vector<U256> elementwise(Context &ctx, int op, const std::vector<U256> &a, const std::vector<U256> &b) {
    std::vector<U256> out(a.size());
    if (a.empty()) return out;
    check(hbmpc_elementwise(ctx.get(), op, a.size(), a[0].data(), b[0].data(), out[0].data()));
    return out;
}
// K5 fused (hbmpc_share_algebra_fused): the algebra of one protocol step in one pass; `in` holds 3 | 4 | 5 vectors of equal length
inline std::vector<std::vector<U256>> fused(Context &ctx, int op, const std::vector<const std::vector<U256> *> &in) {
    const size_t count = in[0]->size(), nout = op == HBMPC_K5_BEAVER_MASK ? 2 : 1;
    std::vector<std::vector<U256>> out(nout, std::vector<U256>(count));
    if (count == 0) return out;
    const uint64_t *pi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t *po[2] = {nullptr, nullptr};
    for (size_t k = 0; k < in.size(); ++k) {
        if (in[k]->size() != count) throw ShareError(HBMPC_INVALID_INPUT);
        pi[k] = (*in[k])[0].data();
    }
    for (size_t k = 0; k < nout; ++k) po[k] = out[k][0].data();
    check(hbmpc_share_algebra_fused(ctx.get(), op, count, pi, po));
    return out;
}
inline std::vector<U256> values(const std::vector<Share> &s) {
    std::vector<U256> v(s.size());
    for (size_t i = 0; i < s.size(); ++i) v[i] = s[i].share;
    return v;
}
// ShamirShare Add / Sub (common/mod.rs:167-213): equal ids and degrees or IdMismatch / DegreeMismatch
inline void same_shape(const std::vector<Share> &a, const std::vector<Share> &b) {
    if (a.size() != b.size()) throw ShareError(HBMPC_INVALID_INPUT);
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].degree != b[i].degree) throw ShareError(HBMPC_DEGREE_MISMATCH);
        if (a[i].id != b[i].id) throw ShareError(HBMPC_ID_MISMATCH);
    }
}
inline std::vector<Share> with_values(const std::vector<Share> &like, const std::vector<U256> &v, size_t degree_add = 0) {
    std::vector<Share> out(like.size());
    for (size_t i = 0; i < like.size(); ++i) out[i] = Share{v[i], like[i].id, like[i].degree + degree_add};
    return out;
}
inline std::vector<U256> deser_secrets(const std::vector<uint8_t> &bytes) {   // Vec<F>::deserialize_compressed of a finished opening
    return deser_bounded_vec(bytes, bytes.size());
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------- TripleGenNode
class TripleGenNode {
   public:
    size_t id, n_parties, threshold;
    BatchReconNode batch_recon_node;                                   // degree 2t (triple_generation.rs:61)
    std::map<SessionId, std::vector<ShamirBeaverTriple>> output;        // output_sender: finished sessions

    TripleGenNode(Context &ctx, size_t id_, size_t n, size_t t) : id(id_), n_parties(n), threshold(t), batch_recon_node(ctx, id_, n, t, 2 * t), ctx_(ctx) {}

    // triple_generation.rs:304-362: flattened groups of 2t+1; sub_share = a*b - r_2t (degree 2t), opened in ONE batched session
    void init_batch(const std::vector<Share> &random_shares_a, const std::vector<Share> &random_shares_b, const std::vector<DoubleShamirShare> &randousha_pairs,
                    SessionId session_id, Network &net) {
        const size_t group_size = 2 * threshold + 1;
        if (randousha_pairs.empty() || randousha_pairs.size() % group_size != 0 || random_shares_a.size() != randousha_pairs.size() ||
            random_shares_b.size() != randousha_pairs.size())
            throw TripleGenError(TripleGenError::NotEnoughPreprocessing, "not enough preprocessing");
        std::vector<Share> r2t(randousha_pairs.size());
        for (size_t i = 0; i < r2t.size(); ++i) r2t[i] = randousha_pairs[i].degree_2t;
        std::vector<Share> sub;
        try {
            // share_mul: ids must match, degrees add (common/mod.rs:280-300); then Sub with the degree-2t share
            for (size_t i = 0; i < random_shares_a.size(); ++i)
                if (random_shares_a[i].id != random_shares_b[i].id) throw ShareError(HBMPC_ID_MISMATCH);
            std::vector<Share> prod = random_shares_a;   // shape of the product share: same ids, degrees added (the values come below)
            for (size_t i = 0; i < prod.size(); ++i) prod[i].degree = random_shares_a[i].degree + random_shares_b[i].degree;
            detail::same_shape(prod, r2t);
            const std::vector<U256> va = detail::values(random_shares_a), vb = detail::values(random_shares_b), vr = detail::values(r2t);
            sub = detail::with_values(prod, detail::fused(ctx_, HBMPC_K5_TRIPLE_MASK, {&va, &vb, &vr})[0]);   // a*b - r_2t in one pass
        } catch (const ShareError &e) {
            throw TripleGenError(TripleGenError::ShareErr, e.what());
        }
        Storage &st = storage_[session_id];
        st.initialized = true;
        st.pairs = randousha_pairs;
        st.a = random_shares_a;
        st.b = random_shares_b;
        if (try_finalize(session_id)) return;   // the opening may have finished before the local init (messages arrive early)
        batch_recon_node.init_batch_reconstruct_many(sub, session_id, net);
    }

    // WrappedMessage::BatchRecon traffic of this node's sessions; finishes the triple session when its opening terminates
    void process(const BatchReconMsg &msg, Network &net) {
        batch_recon_node.batch_recon_handler(msg, net);
        try_finalize(msg.session_id);
    }

   private:
    struct Storage {
        bool initialized = false, finished = false;
        std::vector<DoubleShamirShare> pairs;
        std::vector<Share> a, b;
    };
    Context &ctx_;
    std::map<SessionId, Storage> storage_;

    // triple_generation.rs:165-232: c_i = r_t,i + (a_i*b_i - r_i) for every opened value
    bool try_finalize(SessionId sid) {
        auto it = storage_.find(sid);
        if (it == storage_.end() || !it->second.initialized) return false;
        Storage &st = it->second;
        if (st.finished) return true;
        if (!batch_recon_node.has_secrets(sid)) return false;
        const std::vector<U256> opened = detail::deser_secrets(batch_recon_node.get_store(sid));
        if (opened.size() < st.pairs.size()) return false;
        std::vector<Share> rt(st.pairs.size());
        for (size_t i = 0; i < rt.size(); ++i) rt[i] = st.pairs[i].degree_t;
        std::vector<U256> sub(opened.begin(), opened.begin() + rt.size());
        const std::vector<Share> c = detail::with_values(rt, detail::elementwise(ctx_, 0, detail::values(rt), sub));   // Add<F>: share + public value
        std::vector<ShamirBeaverTriple> triples(rt.size());
        for (size_t i = 0; i < rt.size(); ++i) triples[i] = ShamirBeaverTriple{st.a[i], st.b[i], c[i]};
        st.finished = true;
        output[sid] = std::move(triples);
        return true;
    }
};

// ---------------------------------------------------------------------------------------------------------------- Multiply
// reliable broadcast as far as Multiply uses it: the bytes reach every party (the sender included) as `rbc_deliver(sender, bytes)`
struct Rbc {
    virtual ~Rbc() = default;
    virtual void init(size_t sender, SessionId session_id, const std::vector<uint8_t> &bytes) = 0;
};

class Multiply {
   public:
    size_t id, n, t;
    BatchReconNode batch_recon;                                         // degree t (multiplication.rs:159)
    std::map<SessionId, std::vector<Share>> output;

    Multiply(Context &ctx, size_t id_, size_t n_, size_t t_) : id(id_), n(n_), t(t_), batch_recon(ctx, id_, n_, t_, t_), ctx_(ctx) {}

    // child sessions (multiplication.rs:440-470): a-x values -> slot (exec, 0, 1), b-y values -> slot (exec, 1, 1)
    static SessionId child(SessionId parent, uint8_t dealer, uint8_t round) {
        const unsigned __int128 v = ((unsigned __int128)parent.hi << 64) | parent.lo;
        const uint32_t instance = (uint32_t)v;
        const uint64_t exec = (uint64_t)(v >> 48);   // exec_id sits above sub_id / round_id inside the slot
        return SessionId::make(PROTOCOL_MUL, exec, dealer, round, instance);
    }

    // multiplication.rs:342-484
    void init(SessionId session_id, const std::vector<Share> &x, const std::vector<Share> &y, const std::vector<ShamirBeaverTriple> &beaver_triples, Network &net, Rbc &rbc) {
        if (x.size() != y.size() || x.size() != beaver_triples.size())
            throw MulError(MulError::InvalidInput, "Length of x and y vectors and Beaver triples must match");
        const size_t no_of_mul = x.size(), no_of_batch = no_of_mul / (t + 1), share_len = no_of_mul % (t + 1);
        Storage &st = storage_[session_id];
        st.no_of_mul = no_of_mul;
        st.no_of_batch = no_of_batch;
        st.share_len = share_len;
        st.x = x;
        st.y = y;
        st.mult.resize(no_of_mul);
        std::vector<Share> ta(no_of_mul), tb(no_of_mul);
        for (size_t i = 0; i < no_of_mul; ++i) { st.mult[i] = beaver_triples[i].mult; ta[i] = beaver_triples[i].a; tb[i] = beaver_triples[i].b; }
        st.initialized = true;
        if (share_len == 0) st.openings = std::make_pair(std::vector<U256>{}, std::vector<U256>{});
        if (try_finalize(session_id)) return;
        std::vector<Share> a_sub_x, b_sub_y;
        try {
            detail::same_shape(ta, x);
            detail::same_shape(tb, y);
            const std::vector<U256> va = detail::values(ta), vx = detail::values(x), vb = detail::values(tb), vy = detail::values(y);
            const std::vector<std::vector<U256>> m = detail::fused(ctx_, HBMPC_K5_BEAVER_MASK, {&va, &vx, &vb, &vy});   // a - x, b - y
            a_sub_x = detail::with_values(ta, m[0]);
            b_sub_y = detail::with_values(tb, m[1]);
        } catch (const ShareError &e) {
            throw MulError(MulError::ShareErr, e.what());
        }
        const size_t split_at = no_of_mul - share_len;
        if (split_at > 0) {
            batch_recon.init_batch_reconstruct_many(std::vector<Share>(a_sub_x.begin(), a_sub_x.begin() + split_at), child(session_id, 0, 1), net);
            batch_recon.init_batch_reconstruct_many(std::vector<Share>(b_sub_y.begin(), b_sub_y.begin() + split_at), child(session_id, 1, 1), net);
        }
        if (share_len > 0) {   // ReconstructionMessage(remaining_a, remaining_b) through RBC: share_len values each, 32 bytes per value
            std::vector<uint8_t> bytes(16 + 64 * share_len);
            std::memcpy(bytes.data(), &session_id.lo, 8);
            std::memcpy(bytes.data() + 8, &session_id.hi, 8);
            for (size_t i = 0; i < share_len; ++i) {
                std::memcpy(bytes.data() + 16 + 32 * i, a_sub_x[split_at + i].share.data(), 32);
                std::memcpy(bytes.data() + 16 + 32 * (share_len + i), b_sub_y[split_at + i].share.data(), 32);
            }
            rbc.init(id, child(session_id, (uint8_t)id, 2), bytes);
        }
    }

    // open_mult_handler: BatchRecon traffic of the two child sessions
    void process(const BatchReconMsg &msg, Network &net) {
        batch_recon.batch_recon_handler(msg, net);
        for (auto &kv : storage_) try_finalize(kv.first);
    }
    // RBC output: another party's remainder shares (multiplication.rs: rbc_output handler + reconstruct_rbc once 2t+1 are in)
    void rbc_deliver(size_t sender, const std::vector<uint8_t> &bytes) {
        if (bytes.size() < 16 || (bytes.size() - 16) % 64 != 0) return;   // malformed: ignored like an undecodable message
        SessionId sid;
        std::memcpy(&sid.lo, bytes.data(), 8);
        std::memcpy(&sid.hi, bytes.data() + 8, 8);
        Storage &st = storage_[sid];
        const size_t len = (bytes.size() - 16) / 64;
        std::vector<U256> a(len), b(len);
        for (size_t i = 0; i < len; ++i) {
            std::memcpy(a[i].data(), bytes.data() + 16 + 32 * i, 32);
            std::memcpy(b[i].data(), bytes.data() + 16 + 32 * (len + i), 32);
            if (!fr_is_canonical(a[i]) || !fr_is_canonical(b[i])) return;   // F::deserialize_compressed fails
        }
        st.received[sender] = std::make_pair(std::move(a), std::move(b));
        try_finalize(sid);
    }

   private:
    struct Storage {
        bool initialized = false, finished = false;
        size_t no_of_mul = 0, no_of_batch = 0, share_len = 0;
        std::vector<Share> x, y, mult;
        std::map<size_t, std::pair<std::vector<U256>, std::vector<U256>>> received;   // received_shares by party
        std::optional<std::pair<std::vector<U256>, std::vector<U256>>> openings;       // the remainder values, opened
    };
    Context &ctx_;
    std::map<SessionId, Storage> storage_;

    // reconstruct_rbc (multiplication.rs:101-140): every remainder value from the parties' shares, all 2*share_len codewords in ONE
    // robust-interpolation call (the senders seen so far form the common id set)
    bool reconstruct_rbc(Storage &st) {
        std::vector<size_t> ids;
        for (const auto &kv : st.received)
            if (kv.second.first.size() == st.share_len && kv.second.second.size() == st.share_len) ids.push_back(kv.first);
        const size_t S = ids.size(), B = 2 * st.share_len;
        if (S < 2 * t + 1) return false;
        std::vector<U256> words(B * S), coeffs(B * (t + 1)), secrets(B);
        for (size_t s = 0; s < S; ++s) {
            const auto &pr = st.received[ids[s]];
            for (size_t i = 0; i < st.share_len; ++i) { words[i * S + s] = pr.first[i]; words[(st.share_len + i) * S + s] = pr.second[i]; }
        }
        std::vector<int32_t> path(B);
        const int rc = hbmpc_robust_interpolate_batch(ctx_.get(), n, t, t, S, ids.data(), B, words[0].data(), coeffs[0].data(), secrets[0].data(), path.data(), nullptr);
        if (rc != HBMPC_SUCCESS) return false;   // "could fail if shares corrupt": retried when the next party's message arrives
        st.openings = std::make_pair(std::vector<U256>(secrets.begin(), secrets.begin() + st.share_len), std::vector<U256>(secrets.begin() + st.share_len, secrets.end()));
        return true;
    }

    // finalize_mul (multiplication.rs:57-99): [xy] = [c] - (a-x)(b-y) - (a-x)[y] - (b-y)[x], one K5 call per vector operation
    bool try_finalize(SessionId sid) {
        auto it = storage_.find(sid);
        if (it == storage_.end() || !it->second.initialized) return false;
        Storage &st = it->second;
        if (st.finished) return true;
        const SessionId s1 = child(sid, 0, 1), s2 = child(sid, 1, 1);
        if (st.no_of_batch > 0 && !(batch_recon.has_secrets(s1) && batch_recon.has_secrets(s2))) return false;
        if (!st.openings && !reconstruct_rbc(st)) return false;
        std::vector<U256> da, db;
        if (st.no_of_batch > 0) {
            da = detail::deser_secrets(batch_recon.get_store(s1));
            db = detail::deser_secrets(batch_recon.get_store(s2));
            da.resize(st.no_of_batch * (t + 1));
            db.resize(st.no_of_batch * (t + 1));
        }
        da.insert(da.end(), st.openings->first.begin(), st.openings->first.end());
        db.insert(db.end(), st.openings->second.begin(), st.openings->second.end());
        if (da.size() != st.no_of_mul || db.size() != st.no_of_mul) throw MulError(MulError::InvalidInput, "Inconsistent lengths in finalize_mul");
        // [xy] = [c] - (a-x)(b-y) - (a-x)[y] - (b-y)[x]   (three Mul and three Sub of the reference) in one pass
        const std::vector<U256> vc = detail::values(st.mult), vx = detail::values(st.x), vy = detail::values(st.y);
        const std::vector<U256> z = detail::fused(ctx_, HBMPC_K5_BEAVER_FINALIZE, {&vc, &vx, &vy, &da, &db})[0];
        st.finished = true;
        output[sid] = detail::with_values(st.mult, z);
        return true;
    }
};

}  // namespace hbmpc
