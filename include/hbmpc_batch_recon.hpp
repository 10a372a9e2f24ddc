// hbmpc_batch_recon.hpp -- C++17 host-side mirror of the reference's BatchReconNode (HBMPC Fig. 2, "BatchRecPub") over the
// batch C ABI (hbmpc_b200.h).  SURVEY.md 8(f) N2: the caller either side of the hot path.
//
// The reference host is Rust (not available in this image); this header restates the state machine of
//   mpc/src/honeybadger/batch_recon/batch_recon.rs   (init_batch_reconstruct :103-139, init_batch_reconstruct_many :144-185,
//                                                      batch_recon_handler :191-481, get_or_create_store :491-530)
//   mpc/src/honeybadger/batch_recon/mod.rs            (BatchReconMsgType :18-24, BatchReconMsg :27-33, BatchReconStore :50-58)
// with the reference's names, thresholds and error behaviour, so that tests/host/batch_recon_test.cpp reads like
// mpc/tests/batchrecon_test.rs.  What changes is WHERE the field arithmetic runs: the per-chunk loops over
// apply_vandermonde / batch_recover_secret / RobustShare::recover_secret become ONE device call per message
// (hbmpc_apply_vandermonde_batch with recipient_major = 1, hbmpc_batch_recover_secrets, hbmpc_batch_recover,
// hbmpc_robust_interpolate_batch).  No field arithmetic happens in this header (payload bytes are only copied and
// range-checked); there is no CPU fallback.
//
// Wire formats (recalled from ark-serialize 0.5 / bincode 1.3, not verifiable here):
//   F (compressed)        32 bytes, little-endian canonical value            == the C ABI's U256
//   Vec<F> (compressed)   u64 LE length, then the elements                    -> payload + 8 is a valid host pointer for the C ABI
//   WrappedMessage::BatchRecon(BatchReconMsg) under bincode (fixint, LE): u32 variant index 2 (honeybadger/mod.rs:2168-2177),
//   session_id u128 (16 B), sender_id usize as u64, msg_type u32 variant index, payload u64 length + bytes.
#pragma once
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <optional>

#include "hbmpc_b200.hpp"

namespace hbmpc {

struct SessionId {  // honeybadger/mod.rs:2356 SessionId(u128)
    uint64_t lo = 0, hi = 0;
    bool operator<(const SessionId &o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
    bool operator==(const SessionId &o) const { return hi == o.hi && lo == o.lo; }
    // SessionId::new(protocol, slot, instance_id): instance_id[0..32] slot[32..112] caller[112..120]  (mod.rs:2378-2386)
    static SessionId make(uint8_t protocol, uint64_t exec_id, uint8_t sub_id, uint8_t round_id, uint32_t instance_id) {
        // pack_slot(exec, sub, round): round_id[0..8] sub_id[8..16] exec_id[16..80]
        const unsigned __int128 slot = ((unsigned __int128)exec_id << 16) | ((unsigned __int128)sub_id << 8) | round_id;
        const unsigned __int128 v = ((unsigned __int128)protocol << 112) | (slot << 32) | instance_id;
        return SessionId{(uint64_t)v, (uint64_t)(v >> 64)};
    }
};
inline constexpr uint8_t PROTOCOL_BATCH_RECON = 6;  // ProtocolType::BatchRecon (mod.rs:2197)

enum class BatchReconMsgType : uint32_t { Eval = 0, Reveal = 1, EvalBatch = 2, RevealBatch = 3 };

struct BatchReconMsg {
    SessionId session_id;
    size_t sender_id = 0;
    BatchReconMsgType msg_type = BatchReconMsgType::Eval;
    std::vector<uint8_t> payload;

    static constexpr uint32_t WRAPPED_VARIANT = 2;  // WrappedMessage::BatchRecon
    std::vector<uint8_t> encode() const {           // bincode::serialize(&WrappedMessage::BatchRecon(msg))
        std::vector<uint8_t> out(4 + 16 + 8 + 4 + 8 + payload.size());
        uint8_t *p = out.data();
        auto put = [&p](const void *src, size_t nbytes) { std::memcpy(p, src, nbytes); p += nbytes; };
        const uint32_t tag = WRAPPED_VARIANT, mt = (uint32_t)msg_type;
        const uint64_t sid = (uint64_t)sender_id, len = (uint64_t)payload.size();
        put(&tag, 4); put(&session_id.lo, 8); put(&session_id.hi, 8); put(&sid, 8); put(&mt, 4); put(&len, 8);
        if (!payload.empty()) put(payload.data(), payload.size());
        return out;
    }
    static std::optional<BatchReconMsg> decode(const std::vector<uint8_t> &raw) {  // None: malformed or another variant
        if (raw.size() < 40) return std::nullopt;
        const uint8_t *p = raw.data();
        auto get = [&p](void *dst, size_t nbytes) { std::memcpy(dst, p, nbytes); p += nbytes; };
        uint32_t tag, mt;
        uint64_t sid, len;
        BatchReconMsg m;
        get(&tag, 4); get(&m.session_id.lo, 8); get(&m.session_id.hi, 8); get(&sid, 8); get(&mt, 4); get(&len, 8);
        if (tag != WRAPPED_VARIANT || mt > 3 || len != raw.size() - 40) return std::nullopt;
        m.sender_id = (size_t)sid;
        m.msg_type = (BatchReconMsgType)mt;
        m.payload.assign(p, p + len);
        return m;
    }
};

struct BatchReconError : std::runtime_error {  // batch_recon/mod.rs:75-96
    enum Kind { NetworkError, ShareErr, ArkDeserialization, InvalidInput, InterpolateError, SendError } kind;
    int code;  // ShareErrorCode for ShareErr / InterpolateError
    BatchReconError(Kind k, const std::string &what, int c = 0) : std::runtime_error(what), kind(k), code(c) {}
};

// stoffelnet::network_utils::Network as far as this protocol uses it
struct Network {
    virtual ~Network() = default;
    virtual void send(size_t recipient, const std::vector<uint8_t> &bytes) = 0;
    virtual void broadcast(const std::vector<uint8_t> &bytes) = 0;  // to every party, the sender included
};

struct BatchReconStore {  // batch_recon/mod.rs:50-58
    std::vector<Share> evals_received, reveals_received;
    std::vector<std::pair<size_t, std::vector<U256>>> batch_evals_received, batch_reveals_received;
    std::optional<Share> y_j;
    std::optional<std::vector<U256>> y_j_batch;
    std::optional<std::vector<uint8_t>> secrets;
};

namespace detail {
inline std::vector<uint8_t> ser_vec(const std::vector<U256> &v) {  // Vec<F>::serialize_compressed
    std::vector<uint8_t> out(8 + 32 * v.size());
    const uint64_t len = v.size();
    std::memcpy(out.data(), &len, 8);
    if (!v.empty()) std::memcpy(out.data() + 8, v.data(), 32 * v.size());
    return out;
}
inline U256 deser_f(const uint8_t *p, size_t avail) {  // F::deserialize_compressed: 32 bytes, value < r
    U256 x;
    if (avail < 32) throw BatchReconError(BatchReconError::ArkDeserialization, "short field element");
    std::memcpy(x.data(), p, 32);
    if (!fr_is_canonical(x)) throw BatchReconError(BatchReconError::ArkDeserialization, "non-canonical field element");
    return x;
}
inline std::vector<U256> deser_bounded_vec(const std::vector<uint8_t> &payload, size_t max) {  // common/utils.rs:3-21
    if (payload.size() < 8) throw BatchReconError(BatchReconError::ArkDeserialization, "InvalidData");
    uint64_t len;
    std::memcpy(&len, payload.data(), 8);
    if (len > max) throw BatchReconError(BatchReconError::ArkDeserialization, "InvalidData");
    if (payload.size() - 8 < 32 * len) throw BatchReconError(BatchReconError::ArkDeserialization, "short vector");
    std::vector<U256> v(len);
    for (uint64_t i = 0; i < len; ++i) v[i] = deser_f(payload.data() + 8 + 32 * i, 32);
    return v;
}
}  // namespace detail

class BatchReconNode {
   public:
    size_t id, n, t, degree;
    std::deque<SessionId> output;  // output_sender: sessions whose secrets are ready

    BatchReconNode(Context &ctx, size_t id_, size_t n_, size_t t_, size_t degree_) : id(id_), n(n_), t(t_), degree(degree_), ctx_(ctx) {}

    // batch_recon.rs:103-139: this party's shares of x_0..x_degree -> one Eval message per recipient
    void init_batch_reconstruct(const std::vector<Share> &shares, SessionId session_id, Network &net) {
        if (shares.size() < degree + 1) throw BatchReconError(BatchReconError::InvalidInput, "too little shares to start batch reconstruct");
        std::vector<Share> head(shares.begin(), shares.begin() + degree + 1);
        std::vector<Share> y;
        try {
            y = apply_vandermonde(ctx_, n, head);
        } catch (const ShareError &e) {
            throw BatchReconError(BatchReconError::ShareErr, e.what(), e.code);
        }
        for (size_t j = 0; j < n; ++j) {
            BatchReconMsg m{session_id, id, BatchReconMsgType::Eval, std::vector<uint8_t>(32)};
            std::memcpy(m.payload.data(), y[j].share.data(), 32);
            net.send(j, m.encode());
        }
    }

    // batch_recon.rs:144-185: consecutive chunks of degree+1 secrets; ONE device call encodes every chunk and lays the
    // result out recipient-major, which is the transposition of :158-165 and already the payload order of the wire
    void init_batch_reconstruct_many(const std::vector<Share> &shares, SessionId session_id, Network &net) {
        const size_t width = degree + 1;
        if (shares.empty() || shares.size() % width != 0)
            throw BatchReconError(BatchReconError::InvalidInput, "batched shares must be a non-empty multiple of degree + 1");
        const size_t chunks = shares.size() / width;
        std::vector<U256> in(shares.size());
        for (size_t c = 0; c < chunks; ++c)
            for (size_t k = 0; k < width; ++k) {
                const Share &s = shares[c * width + k], &s0 = shares[c * width];
                if (s.degree != s0.degree) throw BatchReconError(BatchReconError::ShareErr, "DegreeMismatch", HBMPC_DEGREE_MISMATCH);
                if (s.id != s0.id) throw BatchReconError(BatchReconError::ShareErr, "IdMismatch", HBMPC_ID_MISMATCH);
                in[c * width + k] = s.share;
            }
        // the device writes recipient j's vector straight into the payload of the message for j (ark-serialize Vec<F>: u64 length, then
        // the values): the transposition of batch_recon.rs:158-165 and the serialisation of :174-175 are the copy pattern of the call
        std::vector<BatchReconMsg> msgs(n, BatchReconMsg{session_id, id, BatchReconMsgType::EvalBatch, {}});
        std::vector<uint64_t *> out_ptrs(n);
        const uint64_t len = chunks;
        for (size_t j = 0; j < n; ++j) {
            msgs[j].payload.resize(8 + 32 * chunks);
            std::memcpy(msgs[j].payload.data(), &len, 8);
            out_ptrs[j] = reinterpret_cast<uint64_t *>(msgs[j].payload.data() + 8);
        }
        const int rc = hbmpc_apply_vandermonde_msgs(ctx_.get(), n, width, chunks, in[0].data(), out_ptrs.data());
        if (rc != HBMPC_SUCCESS) throw BatchReconError(BatchReconError::ShareErr, "apply_vandermonde", rc);
        for (size_t j = 0; j < n; ++j) net.send(j, msgs[j].encode());
    }

    // batch_recon.rs:191-481
    void batch_recon_handler(const BatchReconMsg &msg, Network &net) {
        const size_t sender_id = msg.sender_id, needed = degree + t + 1;
        switch (msg.msg_type) {
            case BatchReconMsgType::Eval: {
                const U256 val = detail::deser_f(msg.payload.data(), msg.payload.size());
                BatchReconStore *sp = live_store(msg.session_id);
                if (!sp) return;   // terminated session: late / duplicate message, dropped (batch_recon.rs:519-530)
                BatchReconStore &store = *sp;
                if (!seen(store.evals_received, sender_id)) store.evals_received.push_back(Share{val, sender_id, degree});
                if (store.evals_received.size() >= needed && !store.y_j) {
                    U256 value;
                    try {
                        value = RobustShare::recover_secret(ctx_, store.evals_received, n, t).second;
                    } catch (const ShareError &e) {
                        throw BatchReconError(BatchReconError::InterpolateError, e.what(), e.code);
                    }
                    store.y_j = Share{value, id, degree};
                    BatchReconMsg m{msg.session_id, id, BatchReconMsgType::Reveal, std::vector<uint8_t>(32)};
                    std::memcpy(m.payload.data(), value.data(), 32);
                    net.broadcast(m.encode());
                }
                return;
            }
            case BatchReconMsgType::Reveal: {
                const U256 y = detail::deser_f(msg.payload.data(), msg.payload.size());
                BatchReconStore *sp = live_store(msg.session_id);
                if (!sp) return;
                BatchReconStore &store = *sp;
                if (!seen(store.reveals_received, sender_id)) store.reveals_received.push_back(Share{y, sender_id, degree});
                if (store.reveals_received.size() >= needed && !store.secrets) {
                    std::vector<U256> poly;
                    try {
                        poly = RobustShare::recover_secret(ctx_, store.reveals_received, n, t).first;
                    } catch (const ShareError &e) {
                        throw BatchReconError(BatchReconError::InterpolateError, e.what(), e.code);
                    }
                    poly.resize(degree + 1, U256{0, 0, 0, 0});
                    store.secrets = detail::ser_vec(poly);
                    output.push_back(msg.session_id);
                }
                return;
            }
            case BatchReconMsgType::EvalBatch: {
                std::vector<U256> values = detail::deser_bounded_vec(msg.payload, msg.payload.size());
                if (values.empty()) throw BatchReconError(BatchReconError::InvalidInput, "empty EvalBatch payload");
                BatchReconStore *sp = live_store(msg.session_id);
                if (!sp) return;
                BatchReconStore &store = *sp;
                if (!store.batch_evals_received.empty() && store.batch_evals_received[0].second.size() != values.size())
                    throw BatchReconError(BatchReconError::InvalidInput, "inconsistent EvalBatch width");
                if (!seen(store.batch_evals_received, sender_id)) store.batch_evals_received.emplace_back(sender_id, std::move(values));
                if (store.batch_evals_received.size() >= needed && !store.y_j_batch) {
                    // round 1 consumes only P(0) of every chunk (:391): the secrets-only entry point
                    std::vector<U256> y = decode(store.batch_evals_received, true);
                    store.y_j_batch = y;
                    BatchReconMsg m{msg.session_id, id, BatchReconMsgType::RevealBatch, detail::ser_vec(y)};
                    net.broadcast(m.encode());
                }
                return;
            }
            case BatchReconMsgType::RevealBatch: {
                std::vector<U256> values = detail::deser_bounded_vec(msg.payload, msg.payload.size());
                if (values.empty()) throw BatchReconError(BatchReconError::InvalidInput, "empty RevealBatch payload");
                BatchReconStore *sp = live_store(msg.session_id);
                if (!sp) return;
                BatchReconStore &store = *sp;
                if (!store.batch_reveals_received.empty() && store.batch_reveals_received[0].second.size() != values.size())
                    throw BatchReconError(BatchReconError::InvalidInput, "inconsistent RevealBatch width");
                if (!seen(store.batch_reveals_received, sender_id)) store.batch_reveals_received.emplace_back(sender_id, std::move(values));
                if (store.batch_reveals_received.size() >= needed && !store.secrets) {
                    store.secrets = detail::ser_vec(decode(store.batch_reveals_received, false));  // chunk-major, degree+1 each (:463-467)
                    output.push_back(msg.session_id);
                }
                return;
            }
        }
    }
    void process(const BatchReconMsg &msg, Network &net) { batch_recon_handler(msg, net); }

    // batch_recon.rs:81-99
    std::vector<uint8_t> get_store(SessionId session_id) const {
        auto it = store_.find(session_id);
        if (it == store_.end()) throw BatchReconError(BatchReconError::InvalidInput, "Session ID does not exist");
        if (!it->second.secrets) throw BatchReconError(BatchReconError::InvalidInput, "Batch reconstruction has not terminated");
        return *it->second.secrets;
    }
    bool has_secrets(SessionId session_id) const {
        auto it = store_.find(session_id);
        return it != store_.end() && it->second.secrets.has_value();
    }
    bool clear_store(SessionId session_id) { return store_.erase(session_id) > 0; }
    void clear_entire_store() { store_.clear(); }
    size_t store_len() const { return store_.size(); }
    BatchReconStore &get_or_create_store(SessionId session_id) { return store_[session_id]; }
    // get_or_create_store as the handler uses it (batch_recon.rs:491-530): Ok(None) -- here nullptr -- once the session's secrets are
    // set, so that a late or duplicate message of a finished session is dropped instead of being validated and stored again
    BatchReconStore *live_store(SessionId session_id) {
        BatchReconStore &st = store_[session_id];
        return st.secrets ? nullptr : &st;
    }

   private:
    Context &ctx_;
    std::map<SessionId, BatchReconStore> store_;

    static bool seen(const std::vector<Share> &v, size_t sender) {
        for (const Share &s : v)
            if (s.id == sender) return true;
        return false;
    }
    static bool seen(const std::vector<std::pair<size_t, std::vector<U256>>> &v, size_t sender) {
        for (const auto &e : v)
            if (e.first == sender) return true;
        return false;
    }
    // batch_recover_secret(&received, n, degree, t)? : every chunk in one device call.  An Err of the reference (some chunk
    // undecodable with the senders seen so far) leaves the store untouched, so the next message retries with one more sender.
    std::vector<U256> decode(const std::vector<std::pair<size_t, std::vector<U256>>> &received, bool secrets_only) {
        const size_t S = received.size(), B = received[0].second.size();
        std::vector<size_t> ids(S);
        std::vector<const uint64_t *> ptrs(S);   // one array per sender, where its message left it: no gather into [S][B]
        for (size_t i = 0; i < S; ++i) {
            if (received[i].second.size() != B) throw BatchReconError(BatchReconError::InterpolateError, "Inconsistent batch widths", HBMPC_INVALID_INPUT);
            ids[i] = received[i].first;
            ptrs[i] = received[i].second[0].data();
        }
        std::vector<int32_t> path(B);
        std::vector<U256> out(secrets_only ? B : B * (degree + 1));
        const int rc = secrets_only ? hbmpc_batch_recover_secrets_msgs(ctx_.get(), n, degree, t, S, ids.data(), B, ptrs.data(), out[0].data(), path.data())
                                    : hbmpc_batch_recover_msgs(ctx_.get(), n, degree, t, S, ids.data(), B, ptrs.data(), out[0].data(), path.data(), nullptr);
        if (rc != HBMPC_SUCCESS) throw BatchReconError(BatchReconError::InterpolateError, "batch_recover_secret", rc);
        return out;
    }
};

}  // namespace hbmpc
