// hbmpc_ran_dou_sha.hpp -- C++17 host-side mirror of the reference's RanDouShaNode (random double sharing, eprint 2019/883) over
// the batch C ABI (hbmpc_b200.h).  SURVEY.md 8(f) N2/N4: a caller of the hot path and its degree / consistency checks.
//
// Restated from
//   mpc/src/honeybadger/ran_dou_sha/mod.rs        init_batch :371-449, reconstruction_handler :460-617, output_handler :640-668,
//                                                 try_finalize :289-342, RanDouShaStore :75-94
//   mpc/src/honeybadger/ran_dou_sha/messages.rs   RanDouShaPayload :8-13, RanDouShaMessage :16-24, ReconstructionMessage :40-46
// with the reference's thresholds and error behaviour.  What changes is where the field arithmetic runs:
//   init_batch              the two per-batch loops over apply_vandermonde (n x n hyperinvertible matrix) become TWO device calls
//                           (degree t and degree 2t), B = number of batches                    -> hbmpc_apply_vandermonde_batch
//   reconstruction_handler  the per-batch pairs of NonRobustShare::recover_secret + degree checks become two device calls with the
//                           shares sender-major as they arrived                                 -> hbmpc_nonrobust_recover_batch
// The reliable broadcast of the verdict (`self.rbc.init`, common/rbc: out of scope) is a callback here.  No field arithmetic
// happens in this header; there is no CPU fallback.
//
// Wire formats (recalled from ark-serialize 0.5 / bincode 1.3, not verifiable here): a NonRobustShare record is 32-byte LE value +
// u64 id + u64 degree (48 bytes), a ReconstructionMessage two of them (96 bytes); WrappedMessage::RanDouSha is variant 0
// (honeybadger/mod.rs:2168-2177): u32 0, sender_id u64, session_id u128, payload enum u32 {0 Reconstruct(Vec<u8>),
// 1 ReconstructBatch(Vec<Vec<u8>>), 2 Output(bool as u8)}.
#pragma once
#include <functional>

#include "hbmpc_batch_recon.hpp"

namespace hbmpc {

inline constexpr uint8_t PROTOCOL_RANDOUSHA = 1;  // ProtocolType::Randousha (mod.rs:2192)

struct ReconstructionMessage {  // messages.rs:40-46
    Share r_share_deg_t, r_share_deg_2t;
    static void put_share(uint8_t *p, const Share &s) {
        const uint64_t id = s.id, deg = s.degree;
        std::memcpy(p, s.share.data(), 32);
        std::memcpy(p + 32, &id, 8);
        std::memcpy(p + 40, &deg, 8);
    }
    static Share get_share(const uint8_t *p) {
        Share s;
        uint64_t id, deg;
        std::memcpy(s.share.data(), p, 32);
        std::memcpy(&id, p + 32, 8);
        std::memcpy(&deg, p + 40, 8);
        if (!fr_is_canonical(s.share)) throw BatchReconError(BatchReconError::ArkDeserialization, "non-canonical field element");
        s.id = (size_t)id;
        s.degree = (size_t)deg;
        return s;
    }
    std::vector<uint8_t> serialize() const {
        std::vector<uint8_t> out(96);
        put_share(out.data(), r_share_deg_t);
        put_share(out.data() + 48, r_share_deg_2t);
        return out;
    }
    static ReconstructionMessage deserialize(const std::vector<uint8_t> &b) {
        if (b.size() < 96) throw BatchReconError(BatchReconError::ArkDeserialization, "short ReconstructionMessage");
        return ReconstructionMessage{get_share(b.data()), get_share(b.data() + 48)};
    }
};

struct RanDouShaMessage {  // messages.rs:16-24
    enum Kind : uint32_t { Reconstruct = 0, ReconstructBatch = 1, Output = 2 };
    size_t sender_id = 0;
    SessionId session_id;
    Kind kind = Reconstruct;
    std::vector<std::vector<uint8_t>> payloads;  // Reconstruct: one entry; ReconstructBatch: one per batch
    bool ok = false;                             // Output

    static constexpr uint32_t WRAPPED_VARIANT = 0;  // WrappedMessage::RanDouSha
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out;
        auto put = [&out](const void *src, size_t nbytes) { const uint8_t *p = (const uint8_t *)src; out.insert(out.end(), p, p + nbytes); };
        const uint32_t tag = WRAPPED_VARIANT, k = (uint32_t)kind;
        const uint64_t sid = sender_id;
        put(&tag, 4); put(&sid, 8); put(&session_id.lo, 8); put(&session_id.hi, 8); put(&k, 4);
        if (kind == Output) {
            const uint8_t b = ok ? 1 : 0;
            put(&b, 1);
        } else {
            if (kind == ReconstructBatch) { const uint64_t cnt = payloads.size(); put(&cnt, 8); }
            for (const auto &p : payloads) {
                const uint64_t len = p.size();
                put(&len, 8);
                put(p.data(), p.size());
                if (kind == Reconstruct) break;
            }
        }
        return out;
    }
    static std::optional<RanDouShaMessage> decode(const std::vector<uint8_t> &raw) {
        size_t off = 0;
        auto get = [&](void *dst, size_t nbytes) -> bool {
            if (raw.size() - off < nbytes) return false;
            std::memcpy(dst, raw.data() + off, nbytes);
            off += nbytes;
            return true;
        };
        uint32_t tag, k;
        uint64_t sid;
        RanDouShaMessage m;
        if (!get(&tag, 4) || tag != WRAPPED_VARIANT || !get(&sid, 8) || !get(&m.session_id.lo, 8) || !get(&m.session_id.hi, 8) || !get(&k, 4) || k > 2)
            return std::nullopt;
        m.sender_id = (size_t)sid;
        m.kind = (Kind)k;
        if (m.kind == Output) {
            uint8_t b;
            if (!get(&b, 1) || b > 1) return std::nullopt;
            m.ok = b != 0;
        } else {
            uint64_t cnt = 1;
            if (m.kind == ReconstructBatch && (!get(&cnt, 8) || cnt > raw.size())) return std::nullopt;
            for (uint64_t i = 0; i < cnt; ++i) {
                uint64_t len;
                if (!get(&len, 8) || len > raw.size() - off) return std::nullopt;
                m.payloads.emplace_back(raw.begin() + off, raw.begin() + off + len);
                off += len;
            }
        }
        return off == raw.size() ? std::optional<RanDouShaMessage>(m) : std::nullopt;
    }
};

struct RanDouShaError : std::runtime_error {  // ran_dou_sha/mod.rs error enum, as far as this path raises it
    enum Kind { NetworkError, ShareErr, ArkDeserialization, IncorrectID, SessionIdError, Abort } kind;
    int code;
    RanDouShaError(Kind k, const std::string &what, int c = 0) : std::runtime_error(what), kind(k), code(c) {}
};

struct DoubleShamirShare {  // common/share: a degree-t and a degree-2t sharing of the same value
    Share degree_t, degree_2t;
};

struct RanDouShaStore {  // mod.rs:75-94
    std::map<size_t, std::vector<Share>> received_r_shares_degree_t, received_r_shares_degree_2t;
    std::vector<Share> computed_r_shares_degree_t, computed_r_shares_degree_2t;  // [batch][n_parties]
    std::vector<size_t> received_ok_msg;
    size_t batch_size = 0;
    bool finished = false;
    std::vector<DoubleShamirShare> protocol_output;
};

class RanDouShaNode {
   public:
    size_t id, n_parties, threshold;
    // stands in for `self.rbc.init(bytes_msg, sessionid, network)`: reliable broadcast of this checker's verdict
    std::function<void(const RanDouShaMessage &)> broadcast_output;

    RanDouShaNode(Context &ctx, size_t id_, size_t n_, size_t t_) : id(id_), n_parties(n_), threshold(t_), ctx_(ctx) {}

    // mod.rs:371-449.  shares_deg_*_by_batch[b] = this party's shares of the n secrets s_1..s_n of batch b.
    void init_batch(const std::vector<std::vector<Share>> &shares_deg_t_by_batch, const std::vector<std::vector<Share>> &shares_deg_2t_by_batch,
                    SessionId session_id, Network &net) {
        if (sub_id(session_id) != 0) throw RanDouShaError(RanDouShaError::SessionIdError, "sub_id != 0");
        if (shares_deg_t_by_batch.size() != shares_deg_2t_by_batch.size()) throw RanDouShaError(RanDouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        RanDouShaStore &store = store_[session_id];
        store.computed_r_shares_degree_t = hyperinvertible(shares_deg_t_by_batch);
        store.computed_r_shares_degree_2t = hyperinvertible(shares_deg_2t_by_batch);
        store.batch_size = shares_deg_t_by_batch.size();
        if (try_finalize(session_id)) return;
        // party i > t receives the shares of r_i (one pair per batch) and checks them
        for (size_t i = threshold + 1; i < n_parties; ++i) {
            RanDouShaMessage m;
            m.sender_id = id;
            m.session_id = session_id;
            m.kind = store.batch_size == 1 ? RanDouShaMessage::Reconstruct : RanDouShaMessage::ReconstructBatch;
            for (size_t b = 0; b < store.batch_size; ++b)
                m.payloads.push_back(ReconstructionMessage{store.computed_r_shares_degree_t[b * n_parties + i],
                                                           store.computed_r_shares_degree_2t[b * n_parties + i]}.serialize());
            net.send(i, m.encode());
        }
    }

    // mod.rs:460-617.  Returns the verdict when this call ran the check.
    std::optional<bool> reconstruction_handler(const RanDouShaMessage &msg) {
        if (sub_id(msg.session_id) != 0) throw RanDouShaError(RanDouShaError::SessionIdError, "sub_id != 0");
        if (msg.kind == RanDouShaMessage::Output) throw RanDouShaError(RanDouShaError::Abort, "Output payload in the reconstruction phase");
        std::vector<ReconstructionMessage> rec;
        for (const auto &p : msg.payloads) {
            try {
                rec.push_back(ReconstructionMessage::deserialize(p));
            } catch (const BatchReconError &e) {
                throw RanDouShaError(RanDouShaError::ArkDeserialization, e.what());
            }
        }
        const size_t sender_id = msg.sender_id;
        for (const auto &r : rec) {
            if (r.r_share_deg_t.id != sender_id || r.r_share_deg_2t.id != sender_id) throw RanDouShaError(RanDouShaError::IncorrectID, "IncorrectID");
            if (r.r_share_deg_t.degree != threshold || r.r_share_deg_2t.degree != 2 * threshold)
                throw RanDouShaError(RanDouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        }
        RanDouShaStore &store = store_[msg.session_id];
        if (store.received_r_shares_degree_t.empty()) store.batch_size = rec.size();
        else if (store.batch_size != rec.size()) throw RanDouShaError(RanDouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        if (store.finished) return std::nullopt;
        if (store.received_r_shares_degree_t.count(sender_id)) return std::nullopt;  // duplicate: ignored
        std::vector<Share> &vt = store.received_r_shares_degree_t[sender_id], &v2 = store.received_r_shares_degree_2t[sender_id];
        for (const auto &r : rec) { vt.push_back(r.r_share_deg_t); v2.push_back(r.r_share_deg_2t); }
        if (!(id >= threshold + 1 && id < n_parties)) return std::nullopt;
        if (!(store.received_r_shares_degree_t.size() >= 2 * threshold + 1 && store.received_r_shares_degree_2t.size() >= n_parties)) return std::nullopt;
        // every batch column at once: interpolate both sharings from the shares as they arrived (sender-major), require the exact
        // degrees and equal constant terms (mod.rs:573-595)
        const size_t B = store.batch_size, S = store.received_r_shares_degree_t.size();
        std::vector<size_t> ids;
        std::vector<U256> ev_t(S * B), ev_2t(S * B);
        size_t row = 0;
        for (const auto &kv : store.received_r_shares_degree_t) {
            ids.push_back(kv.first);
            const std::vector<Share> &s2 = store.received_r_shares_degree_2t.at(kv.first);
            for (size_t b = 0; b < B; ++b) { ev_t[row * B + b] = kv.second[b].share; ev_2t[row * B + b] = s2[b].share; }
            ++row;
        }
        std::vector<U256> co_t(B * (threshold + 1)), co_2t(B * (2 * threshold + 1)), sec_t(B), sec_2t(B);
        std::vector<int32_t> st_t(B), st_2t(B);
        const int rc1 = hbmpc_nonrobust_recover_batch(ctx_.get(), n_parties, threshold, S, ids.data(), B, ev_t[0].data(), 1, co_t[0].data(), sec_t[0].data(), st_t.data());
        const int rc2 = hbmpc_nonrobust_recover_batch(ctx_.get(), n_parties, 2 * threshold, S, ids.data(), B, ev_2t[0].data(), 1, co_2t[0].data(), sec_2t[0].data(), st_2t.data());
        bool ok = rc1 == HBMPC_SUCCESS && rc2 == HBMPC_SUCCESS;
        for (size_t b = 0; ok && b < B; ++b)
            ok = st_t[b] == (int32_t)threshold && st_2t[b] == (int32_t)(2 * threshold) && sec_t[b] == sec_2t[b];
        RanDouShaMessage out;
        out.sender_id = id;
        out.session_id = msg.session_id;
        out.kind = RanDouShaMessage::Output;
        out.ok = ok;
        if (broadcast_output) broadcast_output(out);
        return ok;
    }

    // mod.rs:640-668
    void output_handler(const RanDouShaMessage &msg) {
        if (msg.kind != RanDouShaMessage::Output) throw RanDouShaError(RanDouShaError::Abort, "not an Output payload");
        if (msg.sender_id < threshold + 1 || msg.sender_id >= n_parties) throw RanDouShaError(RanDouShaError::IncorrectID, "IncorrectID");
        if (!msg.ok) throw RanDouShaError(RanDouShaError::Abort, "Abort");
        RanDouShaStore &store = store_[msg.session_id];
        bool seen = false;
        for (size_t s : store.received_ok_msg) seen = seen || s == msg.sender_id;
        if (!seen) store.received_ok_msg.push_back(msg.sender_id);
        try_finalize(msg.session_id);
    }

    const RanDouShaStore &store(SessionId sid) { return store_[sid]; }

   private:
    Context &ctx_;
    std::map<SessionId, RanDouShaStore> store_;

    static uint8_t sub_id(SessionId s) { return (uint8_t)((s.lo >> 40) & 0xFF); }  // SessionId::sub_id: bits 40..48

    // r = V(n x n) * s for every batch in one device call (make_vandermonde(n, n - 1) + apply_vandermonde per batch, mod.rs:392-403)
    std::vector<Share> hyperinvertible(const std::vector<std::vector<Share>> &by_batch) {
        const size_t B = by_batch.size(), n = n_parties;
        if (B == 0) return {};
        std::vector<U256> in(B * n), out(B * n);
        for (size_t b = 0; b < B; ++b) {
            if (by_batch[b].size() != n) throw RanDouShaError(RanDouShaError::ShareErr, "InvalidInput", HBMPC_INVALID_INPUT);  // row length != shares.len()
            for (size_t k = 0; k < n; ++k) {
                const Share &s = by_batch[b][k], &s0 = by_batch[b][0];
                if (s.degree != s0.degree) throw RanDouShaError(RanDouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
                if (s.id != s0.id) throw RanDouShaError(RanDouShaError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
                in[b * n + k] = s.share;
            }
        }
        const int rc = hbmpc_apply_vandermonde_batch(ctx_.get(), n, n, B, in[0].data(), out[0].data(), 0);
        if (rc != HBMPC_SUCCESS) throw RanDouShaError(RanDouShaError::ShareErr, "apply_vandermonde", rc);
        std::vector<Share> r(B * n);
        for (size_t b = 0; b < B; ++b)
            for (size_t j = 0; j < n; ++j) r[b * n + j] = Share{out[b * n + j], by_batch[b][0].id, by_batch[b][0].degree};
        return r;
    }

    // mod.rs:289-342
    bool try_finalize(SessionId session_id) {
        RanDouShaStore &store = store_[session_id];
        if (store.finished) return true;
        if (store.computed_r_shares_degree_t.size() < store.batch_size * n_parties || store.computed_r_shares_degree_2t.size() < store.batch_size * n_parties)
            return false;
        if (store.batch_size == 0 || store.received_ok_msg.size() < n_parties - (threshold + 1)) return false;
        for (size_t b = 0; b < store.batch_size; ++b)
            for (size_t k = 0; k <= threshold; ++k)
                store.protocol_output.push_back(DoubleShamirShare{store.computed_r_shares_degree_t[b * n_parties + k], store.computed_r_shares_degree_2t[b * n_parties + k]});
        store.finished = true;
        return true;
    }
};

}  // namespace hbmpc
