// hbmpc_ran_sha.hpp -- C++17 host-side mirror of the reference's RanShaNode (random single sharing) over the batch C ABI
// (hbmpc_b200.h).  SURVEY.md 8(f) N2: a caller of the hot path on both sides (share generation K1, hyperinvertible apply K2, robust
// recovery K4).
//
// Restated from
//   mpc/src/honeybadger/share_gen/share_gen.rs   try_finalize :182-217, init_batch :232-289, receive_shares_handler :291-386,
//                                                init_ransha_batch :401-454, reconstruction_handler :456-560, output_handler :562-591
//   mpc/src/honeybadger/share_gen/mod.rs         RanShaStore :61-72, RanShaMessageType :102-109, RanShaPayload :112-120,
//                                                RanShaMessage :124-133
// with the reference's thresholds and error behaviour.  Where the field arithmetic runs:
//   init_batch              `batch_size` calls of RobustShare::compute_shares become ONE device call    -> hbmpc_compute_shares_batch
//                           (the random coefficients are drawn on the host exactly as `DensePolynomial::rand` would, coefficient 0
//                           overwritten by the secret, robust_interpolate.rs:68-69)
//   init_ransha_batch       the per-batch apply_vandermonde (n x n) becomes one device call             -> hbmpc_apply_vandermonde_batch
//   reconstruction_handler  the per-batch RobustShare::recover_secret + degree test becomes one call   -> hbmpc_batch_recover
//                           (sender-major, as the shares arrived; Err of any column == ok = false, share_gen.rs:517-533)
// The reliable broadcast of the verdict (`self.rbc.init`, common/rbc: out of scope) is a callback.  No field arithmetic happens in
// this header; there is no CPU fallback.
//
// Wire formats (recalled from ark-serialize 0.5 / bincode 1.3, not verifiable here): a RobustShare record is 32-byte LE value + u64 id +
// u64 degree (48 bytes), Vec<RobustShare> a u64 count + records; WrappedMessage::RanSha is variant 4 (honeybadger/mod.rs:2168-2177):
// u32 4, sender_id u64, msg_type u32, session_id u128, payload enum u32 {0 Share, 1 SharesBatch, 2 Reconstruct, 3 ReconstructSharesBatch
// (each Vec<u8>: u64 length + bytes), 4 Output(bool as u8)}.
#pragma once
#include "hbmpc_ran_dou_sha.hpp"

namespace hbmpc {

inline constexpr uint8_t PROTOCOL_RANSHA = 2;  // ProtocolType::Ransha (mod.rs:2193)

struct RanShaMessage {  // share_gen/mod.rs:124-133
    enum Type : uint32_t { ShareMessage = 0, ReconstructMessage = 1, OutputMessage = 2 };
    enum Payload : uint32_t { Share = 0, SharesBatch = 1, Reconstruct = 2, ReconstructSharesBatch = 3, Output = 4 };
    size_t sender_id = 0;
    Type msg_type = ShareMessage;
    SessionId session_id;
    Payload kind = Share;
    std::vector<uint8_t> bytes;  // Share / Reconstruct: one 48-byte record; *Batch: Vec<RobustShare>
    bool ok = false;             // Output

    static constexpr uint32_t WRAPPED_VARIANT = 4;  // WrappedMessage::RanSha
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out;
        auto put = [&out](const void *src, size_t nbytes) { const uint8_t *p = (const uint8_t *)src; out.insert(out.end(), p, p + nbytes); };
        const uint32_t tag = WRAPPED_VARIANT, mt = (uint32_t)msg_type, k = (uint32_t)kind;
        const uint64_t sid = sender_id;
        put(&tag, 4); put(&sid, 8); put(&mt, 4); put(&session_id.lo, 8); put(&session_id.hi, 8); put(&k, 4);
        if (kind == Output) {
            const uint8_t b = ok ? 1 : 0;
            put(&b, 1);
        } else {
            const uint64_t len = bytes.size();
            put(&len, 8);
            put(bytes.data(), bytes.size());
        }
        return out;
    }
    static std::optional<RanShaMessage> decode(const std::vector<uint8_t> &raw) {
        size_t off = 0;
        auto get = [&](void *dst, size_t nbytes) -> bool {
            if (raw.size() - off < nbytes) return false;
            std::memcpy(dst, raw.data() + off, nbytes);
            off += nbytes;
            return true;
        };
        uint32_t tag, mt, k;
        uint64_t sid;
        RanShaMessage m;
        if (!get(&tag, 4) || tag != WRAPPED_VARIANT || !get(&sid, 8) || !get(&mt, 4) || mt > 2 || !get(&m.session_id.lo, 8) || !get(&m.session_id.hi, 8) ||
            !get(&k, 4) || k > 4)
            return std::nullopt;
        m.sender_id = (size_t)sid;
        m.msg_type = (Type)mt;
        m.kind = (Payload)k;
        if (m.kind == Output) {
            uint8_t b;
            if (!get(&b, 1) || b > 1) return std::nullopt;
            m.ok = b != 0;
        } else {
            uint64_t len;
            if (!get(&len, 8) || len != raw.size() - off) return std::nullopt;
            m.bytes.assign(raw.begin() + off, raw.end());
            off = raw.size();
        }
        return off == raw.size() ? std::optional<RanShaMessage>(m) : std::nullopt;
    }
};

struct RanShaError : std::runtime_error {  // share_gen/mod.rs:23-57, as far as this path raises it
    enum Kind { NetworkError, ArkDeserialization, ShareErr, SessionIdError, InvalidPartyId, Abort } kind;
    int code;
    RanShaError(Kind k, const std::string &what, int c = 0) : std::runtime_error(what), kind(k), code(c) {}
};

struct RanShaStore {  // share_gen/mod.rs:61-72
    enum State { Initialized, FinishedInitialSharing, Reconstruction, Finished } state = Initialized;
    std::map<size_t, std::vector<Share>> initial_shares, received_r_shares;
    std::vector<bool> reception_tracker;
    std::vector<Share> computed_r_shares;  // [batch][n_parties]
    std::vector<size_t> received_ok_msg;
    size_t batch_size = 0;
    std::vector<Share> protocol_output;
};

class RanShaNode {
   public:
    size_t id, n_parties, threshold;
    std::function<void(const RanShaMessage &)> broadcast_output;  // stands in for `self.rbc.init(...)`

    RanShaNode(Context &ctx, size_t id_, size_t n_, size_t t_) : id(id_), n_parties(n_), threshold(t_), ctx_(ctx) {}

    // share_gen.rs:232-289: deal `batch_size` random secrets, one message per recipient
    void init_batch(SessionId session_id, size_t batch_size, const std::function<uint64_t()> &rng, Network &net) {
        if (sub_id(session_id) != 0) throw RanShaError(RanShaError::SessionIdError, "sub_id != 0");
        batch_size = std::max<size_t>(batch_size, 1);
        if (n_parties <= threshold) throw RanShaError(RanShaError::ShareErr, "InvalidInput", HBMPC_INVALID_INPUT);
        const size_t m = threshold + 1;
        std::vector<U256> coeffs(batch_size * m), shares(batch_size * n_parties);
        for (size_t b = 0; b < batch_size; ++b) {
            const U256 secret = fr_rand(rng);                               // F::rand(rng)
            for (size_t k = 0; k < m; ++k) coeffs[b * m + k] = fr_rand(rng);  // DensePolynomial::rand(degree, rng)
            coeffs[b * m] = secret;                                         // poly[0] = secret
        }
        const int rc = hbmpc_compute_shares_batch(ctx_.get(), n_parties, threshold, batch_size, coeffs[0].data(), shares[0].data());
        if (rc != HBMPC_SUCCESS) throw RanShaError(RanShaError::ShareErr, "compute_shares", rc);
        for (size_t j = 0; j < n_parties; ++j) {
            std::vector<Share> mine(batch_size);
            for (size_t b = 0; b < batch_size; ++b) mine[b] = Share{shares[b * n_parties + j], j, threshold};
            RanShaMessage msg;
            msg.sender_id = id;
            msg.msg_type = RanShaMessage::ShareMessage;
            msg.session_id = session_id;
            msg.kind = batch_size == 1 ? RanShaMessage::Share : RanShaMessage::SharesBatch;
            msg.bytes = batch_size == 1 ? ser_one(mine[0]) : ser_vec(mine);
            net.send(j, msg.encode());
        }
        RanShaStore &st = get_or_create_store(session_id);
        st.batch_size = batch_size;
        st.state = RanShaStore::Initialized;
    }

    // share_gen.rs:291-386
    void receive_shares_handler(const RanShaMessage &msg, Network &net) {
        if (sub_id(msg.session_id) != 0) throw RanShaError(RanShaError::SessionIdError, "sub_id != 0");
        if (msg.sender_id >= n_parties) throw RanShaError(RanShaError::InvalidPartyId, "InvalidPartyId");
        if (msg.kind != RanShaMessage::Share && msg.kind != RanShaMessage::SharesBatch) throw RanShaError(RanShaError::Abort, "Abort");
        const std::vector<Share> shares = msg.kind == RanShaMessage::Share ? std::vector<Share>{deser_one(msg.bytes)} : deser_vec(msg.bytes);
        for (const Share &s : shares) {
            if (s.id != id) throw RanShaError(RanShaError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
            if (s.degree != threshold) throw RanShaError(RanShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        }
        RanShaStore &st = get_or_create_store(msg.session_id);
        if (st.initial_shares.empty()) st.batch_size = shares.size();
        else if (st.batch_size != shares.size()) throw RanShaError(RanShaError::Abort, "Abort");
        if (st.state == RanShaStore::FinishedInitialSharing || st.state == RanShaStore::Finished) return;
        if (st.initial_shares.count(msg.sender_id)) return;  // duplicate: ignored
        st.initial_shares[msg.sender_id] = shares;
        st.reception_tracker[msg.sender_id] = true;
        for (bool got : st.reception_tracker)
            if (!got) return;
        st.state = RanShaStore::FinishedInitialSharing;
        std::vector<std::vector<Share>> by_batch(st.batch_size);
        for (const auto &kv : st.initial_shares)  // sorted by sender id
            for (size_t b = 0; b < st.batch_size; ++b) by_batch[b].push_back(kv.second[b]);
        init_ransha_batch(by_batch, msg.session_id, net);
    }

    // share_gen.rs:401-454
    void init_ransha_batch(const std::vector<std::vector<Share>> &shares_by_batch, SessionId session_id, Network &net) {
        const size_t B = shares_by_batch.size(), n = n_parties;
        std::vector<U256> in(B * n), out(B * n);
        for (size_t b = 0; b < B; ++b) {
            if (shares_by_batch[b].size() != n) throw RanShaError(RanShaError::ShareErr, "InvalidInput", HBMPC_INVALID_INPUT);
            for (size_t k = 0; k < n; ++k) {
                const Share &s = shares_by_batch[b][k], &s0 = shares_by_batch[b][0];
                if (s.degree != s0.degree) throw RanShaError(RanShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
                if (s.id != s0.id) throw RanShaError(RanShaError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
                in[b * n + k] = s.share;
            }
        }
        const int rc = hbmpc_apply_vandermonde_batch(ctx_.get(), n, n, B, in[0].data(), out[0].data(), 0);
        if (rc != HBMPC_SUCCESS) throw RanShaError(RanShaError::ShareErr, "apply_vandermonde", rc);
        RanShaStore &st = get_or_create_store(session_id);
        st.batch_size = B;
        st.computed_r_shares.resize(B * n);
        for (size_t b = 0; b < B; ++b)
            for (size_t j = 0; j < n; ++j) st.computed_r_shares[b * n + j] = Share{out[b * n + j], shares_by_batch[b][0].id, shares_by_batch[b][0].degree};
        if (try_finalize(session_id)) return;
        for (size_t i = 0; i < 2 * threshold; ++i) {  // the first 2t parties verify r_i
            std::vector<Share> mine(B);
            for (size_t b = 0; b < B; ++b) mine[b] = st.computed_r_shares[b * n + i];
            RanShaMessage msg;
            msg.sender_id = id;
            msg.msg_type = RanShaMessage::ReconstructMessage;
            msg.session_id = session_id;
            msg.kind = B == 1 ? RanShaMessage::Reconstruct : RanShaMessage::ReconstructSharesBatch;
            msg.bytes = B == 1 ? ser_one(mine[0]) : ser_vec(mine);
            net.send(i, msg.encode());
        }
    }

    // share_gen.rs:456-560.  Returns the verdict when this call ran the check (every message from the (2t+1)-th on does).
    std::optional<bool> reconstruction_handler(const RanShaMessage &msg) {
        if (sub_id(msg.session_id) != 0) throw RanShaError(RanShaError::SessionIdError, "sub_id != 0");
        if (msg.kind != RanShaMessage::Reconstruct && msg.kind != RanShaMessage::ReconstructSharesBatch) throw RanShaError(RanShaError::Abort, "Abort");
        const std::vector<Share> shares = msg.kind == RanShaMessage::Reconstruct ? std::vector<Share>{deser_one(msg.bytes)} : deser_vec(msg.bytes);
        for (const Share &s : shares) {
            if (s.degree != threshold) throw RanShaError(RanShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
            if (s.id != msg.sender_id) throw RanShaError(RanShaError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
        }
        RanShaStore &st = get_or_create_store(msg.session_id);
        if (st.state == RanShaStore::Finished) return std::nullopt;
        if (st.received_r_shares.empty()) st.batch_size = shares.size();
        else if (st.batch_size != shares.size()) throw RanShaError(RanShaError::Abort, "Abort");
        st.state = RanShaStore::Reconstruction;
        st.received_r_shares[msg.sender_id] = shares;
        if (!(id < 2 * threshold && st.received_r_shares.size() >= 2 * threshold + 1)) return std::nullopt;
        // every batch column in one call: robust recovery from the shares as they arrived, then "degree == t" (coefficient t != 0)
        const size_t B = st.batch_size, S = st.received_r_shares.size(), m = threshold + 1;
        std::vector<size_t> ids;
        std::vector<U256> evals(S * B), coeffs(B * m);
        std::vector<int32_t> path(B);
        size_t row = 0;
        for (const auto &kv : st.received_r_shares) {
            ids.push_back(kv.first);
            for (size_t b = 0; b < B; ++b) evals[row * B + b] = kv.second[b].share;
            ++row;
        }
        const int rc = hbmpc_batch_recover(ctx_.get(), n_parties, threshold, threshold, S, ids.data(), B, evals[0].data(), coeffs[0].data(), path.data(), nullptr);
        bool ok = rc == HBMPC_SUCCESS;
        for (size_t b = 0; ok && b < B; ++b) ok = !(coeffs[b * m + threshold] == U256{0, 0, 0, 0}) || threshold == 0;
        RanShaMessage out;
        out.sender_id = id;
        out.msg_type = RanShaMessage::OutputMessage;
        out.session_id = msg.session_id;
        out.kind = RanShaMessage::Output;
        out.ok = ok;
        if (broadcast_output) broadcast_output(out);
        return ok;
    }

    // share_gen.rs:562-591
    void output_handler(const RanShaMessage &msg) {
        if (msg.kind != RanShaMessage::Output) throw RanShaError(RanShaError::Abort, "Abort");
        if (!msg.ok) throw RanShaError(RanShaError::Abort, "Abort");
        if (sub_id(msg.session_id) != 0) throw RanShaError(RanShaError::SessionIdError, "sub_id != 0");
        if (msg.sender_id >= 2 * threshold) throw RanShaError(RanShaError::InvalidPartyId, "InvalidPartyId");
        RanShaStore &st = get_or_create_store(msg.session_id);
        bool seen = false;
        for (size_t s : st.received_ok_msg) seen = seen || s == msg.sender_id;
        if (!seen) st.received_ok_msg.push_back(msg.sender_id);
        try_finalize(msg.session_id);
    }

    // share_gen.rs:593-: dispatch on the message type
    void process(const RanShaMessage &msg, Network &net) {
        switch (msg.msg_type) {
            case RanShaMessage::ShareMessage: receive_shares_handler(msg, net); break;
            case RanShaMessage::ReconstructMessage: reconstruction_handler(msg); break;
            case RanShaMessage::OutputMessage: output_handler(msg); break;
        }
    }

    RanShaStore &get_or_create_store(SessionId sid) {
        auto it = store_.find(sid);
        if (it == store_.end()) {
            it = store_.emplace(sid, RanShaStore{}).first;
            it->second.reception_tracker.assign(n_parties, false);
        }
        return it->second;
    }

   private:
    Context &ctx_;
    std::map<SessionId, RanShaStore> store_;

    static uint8_t sub_id(SessionId s) { return (uint8_t)((s.lo >> 40) & 0xFF); }
    static std::vector<uint8_t> ser_one(const Share &s) {
        std::vector<uint8_t> out(48);
        ReconstructionMessage::put_share(out.data(), s);
        return out;
    }
    static std::vector<uint8_t> ser_vec(const std::vector<Share> &v) {
        std::vector<uint8_t> out(8 + 48 * v.size());
        const uint64_t len = v.size();
        std::memcpy(out.data(), &len, 8);
        for (size_t i = 0; i < v.size(); ++i) ReconstructionMessage::put_share(out.data() + 8 + 48 * i, v[i]);
        return out;
    }
    static Share deser_one(const std::vector<uint8_t> &b) {
        if (b.size() < 48) throw RanShaError(RanShaError::ArkDeserialization, "short share record");
        try {
            return ReconstructionMessage::get_share(b.data());
        } catch (const BatchReconError &e) {
            throw RanShaError(RanShaError::ArkDeserialization, e.what());
        }
    }
    static std::vector<Share> deser_vec(const std::vector<uint8_t> &b) {  // deser_bounded_vec(payload, payload.len())
        if (b.size() < 8) throw RanShaError(RanShaError::ArkDeserialization, "InvalidData");
        uint64_t len;
        std::memcpy(&len, b.data(), 8);
        if (len > b.size() || b.size() - 8 < 48 * len) throw RanShaError(RanShaError::ArkDeserialization, "InvalidData");
        std::vector<Share> v(len);
        try {
            for (uint64_t i = 0; i < len; ++i) v[i] = ReconstructionMessage::get_share(b.data() + 8 + 48 * i);
        } catch (const BatchReconError &e) {
            throw RanShaError(RanShaError::ArkDeserialization, e.what());
        }
        return v;
    }

    // share_gen.rs:182-217
    bool try_finalize(SessionId session_id) {
        RanShaStore &st = get_or_create_store(session_id);
        if (st.state == RanShaStore::Finished) return true;
        if (st.received_ok_msg.size() < 2 * threshold) return false;
        if (st.batch_size == 0 || st.computed_r_shares.size() < st.batch_size * n_parties) return false;
        for (size_t b = 0; b < st.batch_size; ++b)
            for (size_t k = 2 * threshold; k < n_parties; ++k) st.protocol_output.push_back(st.computed_r_shares[b * n_parties + k]);
        st.state = RanShaStore::Finished;
        return true;
    }
};

}  // namespace hbmpc
