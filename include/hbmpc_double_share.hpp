// hbmpc_double_share.hpp -- C++17 host-side mirror of the reference's DoubleShareNode (faulty double-share distribution, the input of
// RanDouSha) over the batch C ABI (hbmpc_b200.h).  SURVEY.md 8(f) N2.
//
// Restated from mpc/src/honeybadger/double_share/double_share_generation.rs (init_batch :151-215, receive_double_shares_handler
// :217-300) and double_share/mod.rs (DoubleShamirShare :60-66, DouShaPayload :80-83, DouShaMessage :87-94).  `batch_size` pairs of
// NonRobustShare::compute_shares calls (degree t and degree 2t of the same secret) become TWO device calls
// (hbmpc_compute_shares_batch with B = batch_size); the coefficients are drawn on the host in the reference's order (secret, the t+1
// coefficients of the degree-t polynomial, the 2t+1 of the degree-2t one; coefficient 0 of both overwritten by the secret).
//
// Wire formats (recalled, not verifiable here): DoubleShamirShare = degree_2t record then degree_t record (48 bytes each, field order of
// the struct); WrappedMessage::Dousha is variant 5: u32 5, sender_id u64, session_id u128, payload enum u32 {0 Share, 1 Shares} + Vec<u8>.
#pragma once
#include "hbmpc_ran_dou_sha.hpp"

namespace hbmpc {

inline constexpr uint8_t PROTOCOL_DOUSHA = 7;  // ProtocolType::Dousha (mod.rs:2198)

struct DouShaMessage {  // double_share/mod.rs:87-94
    enum Payload : uint32_t { Share = 0, Shares = 1 };
    size_t sender_id = 0;
    SessionId session_id;
    Payload kind = Share;
    std::vector<uint8_t> bytes;

    static constexpr uint32_t WRAPPED_VARIANT = 5;  // WrappedMessage::Dousha
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out;
        auto put = [&out](const void *src, size_t nbytes) { const uint8_t *p = (const uint8_t *)src; out.insert(out.end(), p, p + nbytes); };
        const uint32_t tag = WRAPPED_VARIANT, k = (uint32_t)kind;
        const uint64_t sid = sender_id, len = bytes.size();
        put(&tag, 4); put(&sid, 8); put(&session_id.lo, 8); put(&session_id.hi, 8); put(&k, 4); put(&len, 8);
        put(bytes.data(), bytes.size());
        return out;
    }
    static std::optional<DouShaMessage> decode(const std::vector<uint8_t> &raw) {
        if (raw.size() < 40) return std::nullopt;
        const uint8_t *p = raw.data();
        auto get = [&p](void *dst, size_t nbytes) { std::memcpy(dst, p, nbytes); p += nbytes; };
        uint32_t tag, k;
        uint64_t sid, len;
        DouShaMessage m;
        get(&tag, 4); get(&sid, 8); get(&m.session_id.lo, 8); get(&m.session_id.hi, 8); get(&k, 4); get(&len, 8);
        if (tag != WRAPPED_VARIANT || k > 1 || len != raw.size() - 40) return std::nullopt;
        m.sender_id = (size_t)sid;
        m.kind = (Payload)k;
        m.bytes.assign(p, p + len);
        return m;
    }
};

struct DouShaError : std::runtime_error {
    enum Kind { ArkDeserialization, ShareErr, InvalidPartyId } kind;
    int code;
    DouShaError(Kind k, const std::string &what, int c = 0) : std::runtime_error(what), kind(k), code(c) {}
};

struct DouShaStore {
    std::map<size_t, std::vector<DoubleShamirShare>> share;  // sender -> its double shares for this party, one per batch
    std::vector<bool> reception_tracker;
    size_t batch_size = 0;
    bool finished = false;
    std::vector<DoubleShamirShare> protocol_output;  // [batch][sender]
};

class DoubleShareNode {
   public:
    size_t id, n_parties, threshold;
    DoubleShareNode(Context &ctx, size_t id_, size_t n_, size_t t_) : id(id_), n_parties(n_), threshold(t_), ctx_(ctx) {}

    // double_share_generation.rs:151-215
    void init_batch(SessionId session_id, size_t batch_size, const std::function<uint64_t()> &rng, Network &net) {
        batch_size = std::max<size_t>(batch_size, 1);
        const size_t t = threshold, n = n_parties, m1 = t + 1, m2 = 2 * t + 1;
        if (n <= 2 * t) throw DouShaError(DouShaError::ShareErr, "InvalidInput", HBMPC_INVALID_INPUT);  // NonRobustShare::compute_shares: n <= degree
        std::vector<U256> c1(batch_size * m1), c2(batch_size * m2), s1(batch_size * n), s2(batch_size * n);
        for (size_t b = 0; b < batch_size; ++b) {
            const U256 secret = fr_rand(rng);
            for (size_t k = 0; k < m1; ++k) c1[b * m1 + k] = fr_rand(rng);
            c1[b * m1] = secret;
            for (size_t k = 0; k < m2; ++k) c2[b * m2 + k] = fr_rand(rng);
            c2[b * m2] = secret;
        }
        int rc = hbmpc_compute_shares_batch(ctx_.get(), n, t, batch_size, c1[0].data(), s1[0].data());
        if (rc == HBMPC_SUCCESS) rc = hbmpc_compute_shares_batch(ctx_.get(), n, 2 * t, batch_size, c2[0].data(), s2[0].data());
        if (rc != HBMPC_SUCCESS) throw DouShaError(DouShaError::ShareErr, "compute_shares", rc);
        for (size_t j = 0; j < n; ++j) {
            DouShaMessage msg;
            msg.sender_id = id;
            msg.session_id = session_id;
            msg.kind = batch_size == 1 ? DouShaMessage::Share : DouShaMessage::Shares;
            msg.bytes.assign((batch_size == 1 ? 0 : 8) + 96 * batch_size, 0);
            size_t off = 0;
            if (batch_size > 1) { const uint64_t len = batch_size; std::memcpy(msg.bytes.data(), &len, 8); off = 8; }
            for (size_t b = 0; b < batch_size; ++b) {
                ReconstructionMessage::put_share(msg.bytes.data() + off + 96 * b, Share{s2[b * n + j], j, 2 * t});       // degree_2t first
                ReconstructionMessage::put_share(msg.bytes.data() + off + 96 * b + 48, Share{s1[b * n + j], j, t});
            }
            net.send(j, msg.encode());
        }
        DouShaStore &st = get_or_create_store(session_id);
        st.batch_size = batch_size;
    }

    // double_share_generation.rs:217-300.  Returns true when every party's double shares have arrived (protocol_output is ready).
    bool receive_double_shares_handler(const DouShaMessage &msg) {
        size_t off = 0, count = 1;
        if (msg.kind == DouShaMessage::Shares) {
            if (msg.bytes.size() < 8) throw DouShaError(DouShaError::ArkDeserialization, "InvalidData");
            uint64_t len;
            std::memcpy(&len, msg.bytes.data(), 8);
            if (len > msg.bytes.size()) throw DouShaError(DouShaError::ArkDeserialization, "InvalidData");
            count = (size_t)len;
            off = 8;
        }
        if (msg.bytes.size() - off < 96 * count) throw DouShaError(DouShaError::ArkDeserialization, "short double share");
        std::vector<DoubleShamirShare> ds(count);
        try {
            for (size_t b = 0; b < count; ++b) {
                ds[b].degree_2t = ReconstructionMessage::get_share(msg.bytes.data() + off + 96 * b);
                ds[b].degree_t = ReconstructionMessage::get_share(msg.bytes.data() + off + 96 * b + 48);
            }
        } catch (const BatchReconError &e) {
            throw DouShaError(DouShaError::ArkDeserialization, e.what());
        }
        for (const DoubleShamirShare &d : ds) {
            if (d.degree_t.id != id || d.degree_2t.id != id) throw DouShaError(DouShaError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
            if (d.degree_t.degree != threshold || d.degree_2t.degree != 2 * threshold) throw DouShaError(DouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        }
        DouShaStore &st = get_or_create_store(msg.session_id);
        if (st.share.empty()) st.batch_size = ds.size();
        else if (st.batch_size != ds.size()) throw DouShaError(DouShaError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
        if (st.finished) return true;
        if (st.share.count(msg.sender_id)) return false;  // duplicate: ignored
        if (msg.sender_id >= n_parties) throw DouShaError(DouShaError::InvalidPartyId, "InvalidPartyId");
        st.share[msg.sender_id] = ds;
        st.reception_tracker[msg.sender_id] = true;
        for (bool got : st.reception_tracker)
            if (!got) return false;
        for (size_t b = 0; b < st.batch_size; ++b)
            for (const auto &kv : st.share) st.protocol_output.push_back(kv.second[b]);
        st.finished = true;
        return true;
    }

    DouShaStore &get_or_create_store(SessionId sid) {
        auto it = store_.find(sid);
        if (it == store_.end()) {
            it = store_.emplace(sid, DouShaStore{}).first;
            it->second.reception_tracker.assign(n_parties, false);
        }
        return it->second;
    }

   private:
    Context &ctx_;
    std::map<SessionId, DouShaStore> store_;
};

}  // namespace hbmpc
