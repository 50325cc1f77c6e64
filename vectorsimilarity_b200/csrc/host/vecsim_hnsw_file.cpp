// HNSW index files in the reference's serialized format: a reader for encoding V3 and V4 that lands
// the rows in the device store and the links in the device graph, and a V4 writer. Format as the reference writes and
// restores it:
//   header              /root/reference/src/VecSim/index_factories/hnsw_factory.cpp:171-205 (version, algo, dim, type,
//                       metric, blockSize, multi, initial capacity)
//   index fields        algorithms/hnsw/hnsw_serializer_impl.h:145-166 restoreIndexFields / :247-276 saveIndexFields
//   metadata + graph    hnsw_serializer_impl.h:168-245 restoreGraph / restoreLevel, :278-330 saveGraph / saveLevel
//   vector blocks       containers/data_blocks_container.cpp:65-110 (V3 stores block count and lengths, V4 does not)
// The traversal never reads the "incoming unidirectional edges" lists, so the reader skips them; the writer derives them
// from the links (u -> v without v -> u) so that the reference can load the file and pass its integrity check.
#include "vecsim_index.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

namespace vsb {

namespace {
struct Reader {
    std::vector<char> buf;
    size_t pos = 0;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if (pos + sizeof(T) > buf.size()) {
            ok = false;
            return v;
        }
        std::memcpy(&v, buf.data() + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    const char *bytes(size_t n) {
        if (pos + n > buf.size()) {
            ok = false;
            return nullptr;
        }
        const char *p = buf.data() + pos;
        pos += n;
        return p;
    }
};
constexpr uint8_t DELETE_MARK = 0x1; // hnsw.h:60
template <typename T> void put(std::ofstream &o, const T &v) { o.write((const char *)&v, sizeof(T)); }
} // namespace

VecSimIndexInterface *load_hnsw_file(const char *path, std::string &err) {
    Reader r;
    {
        std::ifstream in(path, std::ios::binary | std::ios::ate);
        if (!in.is_open()) {
            err = "Cannot open file";
            return nullptr;
        }
        const std::streamsize sz = in.tellg();
        in.seekg(0);
        r.buf.resize((size_t)sz);
        if (sz > 0 && !in.read(r.buf.data(), sz)) {
            err = "Cannot read file";
            return nullptr;
        }
    }
    const int version = r.get<int>();
    if (!r.ok || version <= 2) {
        err = "Cannot load index: deprecated encoding version: " + std::to_string(version);
        return nullptr;
    }
    if (version >= 5) {
        err = "Cannot load index: bad encoding version: " + std::to_string(version);
        return nullptr;
    }
    const int algo = r.get<int>();
    if (algo != VecSimAlgo_HNSWLIB) {
        err = "Cannot load index: Expected HNSW file but got algorithm type: " + std::to_string(algo);
        return nullptr;
    }
    HNSWParams p{};
    p.dim = r.get<size_t>();
    p.type = (VecSimType)r.get<int>();
    p.metric = (VecSimMetric)r.get<int>();
    p.blockSize = r.get<size_t>();
    p.multi = r.get<bool>();
    p.initialCapacity = r.get<size_t>();
    p.M = r.get<size_t>();
    const size_t M0 = r.get<size_t>();
    p.efConstruction = r.get<size_t>();
    p.efRuntime = r.get<size_t>();
    p.epsilon = r.get<double>();
    (void)r.get<double>(); // mult = 1 / ln(M): recomputed by the index
    const size_t n = r.get<size_t>();
    // every element costs at least its label + flag byte + one stored row + one level-0 record header in the file: a
    // count that cannot fit is a corrupted header (and would overflow the buffer sizes computed from it)
    if (r.ok && p.dim && p.dim <= (1u << 24) && n > r.buf.size() / (sizeof(size_t) + 1 + p.dim)) {
        err = "Cannot load index: corrupted header (element count exceeds the file size)";
        return nullptr;
    }
    const size_t num_deleted = r.get<size_t>();
    const size_t max_level = r.get<size_t>();
    const idType entry = r.get<idType>();
    if (!r.ok || p.dim == 0 || p.dim > (1u << 24) || p.type > VecSimType_UINT8 || p.metric > VecSimMetric_Cosine ||
        M0 != 2 * p.M || p.M < 2 || p.M > 256) {
        err = "Cannot load index: corrupted header";
        return nullptr;
    }
    std::vector<size_t> labels(n);
    std::vector<uint8_t> flags(n);
    for (size_t i = 0; i < n && r.ok; i++) {
        labels[i] = r.get<size_t>();
        flags[i] = r.get<uint8_t>();
    }
    const size_t stored = stored_size(p.type, p.dim, p.metric);
    std::vector<uint8_t> rows(n * stored);
    size_t num_blocks;
    if (version == 3) {
        num_blocks = r.get<unsigned int>();
        size_t got = 0;
        for (size_t b = 0; b < num_blocks && r.ok; b++) {
            const size_t len = r.get<unsigned int>();
            if (got + len > n) {
                r.ok = false;
                break;
            }
            const char *src = r.bytes(len * stored);
            if (src) std::memcpy(rows.data() + got * stored, src, len * stored);
            got += len;
        }
        if (got != n) r.ok = false;
    } else {
        num_blocks = p.blockSize ? (size_t)std::ceil((float)n / (float)p.blockSize) : 0;
        const char *src = r.bytes(n * stored);
        if (src) std::memcpy(rows.data(), src, n * stored);
    }
    if (!r.ok) {
        err = "Cannot load index: truncated file (metadata / vectors)";
        return nullptr;
    }
    const size_t w0 = 2 * p.M + 1, wu = p.M + 1;
    std::vector<uint32_t> levels(n), l0(n * w0, 0), upper;
    size_t id = 0;
    for (size_t b = 0; b < num_blocks && r.ok; b++) {
        const size_t len = r.get<unsigned int>();
        for (size_t j = 0; j < len && r.ok; j++, id++) {
            if (id >= n) {
                r.ok = false;
                break;
            }
            const size_t top = r.get<size_t>();
            levels[id] = (uint32_t)top;
            for (size_t lvl = 0; lvl <= top && r.ok; lvl++) {
                const size_t cnt = r.get<uint16_t>();
                uint32_t *rec;
                if (lvl == 0) rec = l0.data() + id * w0;
                else {
                    upper.resize(upper.size() + wu, 0);
                    rec = upper.data() + upper.size() - wu;
                }
                if (cnt > (lvl == 0 ? 2 * p.M : p.M)) {
                    r.ok = false;
                    break;
                }
                rec[0] = (uint32_t)cnt;
                const char *src = r.bytes(cnt * sizeof(idType));
                if (src) std::memcpy(rec + 1, src, cnt * sizeof(idType));
                const size_t incoming = r.get<unsigned int>();
                r.bytes(incoming * sizeof(idType)); // not needed by any traversal
            }
        }
    }
    if (!r.ok || id != n) {
        err = "Cannot load index: truncated or corrupted graph section";
        return nullptr;
    }
    // the traversal kernels follow these ids without bounds checks: refuse anything that points outside the graph
    {
        bool sane = true;
        size_t u = 0;
        for (size_t i = 0; i < n && sane; i++) {
            const uint32_t *rec = l0.data() + i * w0;
            for (uint32_t j = 0; j < rec[0] && sane; j++) sane = rec[1 + j] < n;
            for (uint32_t l = 1; l <= levels[i] && sane; l++, u++) {
                const uint32_t *ur = upper.data() + u * wu;
                for (uint32_t j = 0; j < ur[0] && sane; j++) sane = ur[1 + j] < n && levels[ur[1 + j]] >= l;
            }
        }
        if (n) {
            const bool empty_entry = entry == (idType)-1 || max_level == (size_t)-1;
            if (empty_entry) sane = false;
            else sane = sane && entry < n && levels[entry] == max_level;
        }
        if (!sane) {
            err = "Cannot load index: corrupted graph (link, entry point or level out of range)";
            return nullptr;
        }
    }
    p.initialCapacity = std::max(p.initialCapacity, n);
    auto *idx = new HnswIndex(p, nullptr);
    if (!idx->ok()) {
        err = std::string("HNSW index: ") + vsgpu_last_error();
        delete idx;
        return nullptr;
    }
    const long e = (n == 0 || entry == (idType)-1) ? -1 : (long)entry;
    const long ml = (n == 0 || max_level == (size_t)-1) ? -1 : (long)max_level;
    if (n && idx->importGraph(rows.data(), 1, n, labels.data(), levels.data(), l0.data(), upper.empty() ? nullptr : upper.data(),
                              upper.size() / wu, e, ml) != 0) {
        err = std::string("HNSW import: ") + vsgpu_last_error();
        delete idx;
        return nullptr;
    }
    size_t marked = 0;
    for (size_t i = 0; i < n; i++)
        if (flags[i] & DELETE_MARK) {
            if (idx->markDeletedById((idType)i) != 0) {
                err = "HNSW import: cannot restore a deleted mark";
                delete idx;
                return nullptr;
            }
            marked++;
        }
    if (marked != num_deleted) {
        err = "Cannot load index: deleted marks do not match the header count";
        delete idx;
        return nullptr;
    }
    return idx;
}

// Restores one DELETE_MARK of a loaded file. The node is always flagged on the device (a freshly imported graph has no
// flags at all); the label mapping goes only when it still points at this node — a label that was overwritten or re-added
// after the delete has a later, live id that importGraph's last-id-wins pass already mapped it to.
int HnswIndex::markDeletedById(idType id) {
    std::lock_guard<std::mutex> g(mu_);
    if (id >= id_to_label_.size()) return -1;
    if (markDeletedLocked(id) != 0) return -1;
    if (multi_) {
        auto it = label_to_ids_.find(id_to_label_[id]);
        if (it != label_to_ids_.end()) {
            auto &v = it->second;
            v.erase(std::remove(v.begin(), v.end(), id), v.end());
            if (v.empty()) label_to_ids_.erase(it);
        }
        return 0;
    }
    auto it = label_to_id_.find(id_to_label_[id]);
    if (it != label_to_id_.end() && it->second == id) label_to_id_.erase(it);
    return 0;
}

// HNSWIndex::saveIndex, encoding V4 (hnsw_serializer.cpp:39-52, hnsw_serializer_impl.h:247-330)
int HnswIndex::saveFile(const char *path) {
    std::lock_guard<std::mutex> g(mu_);
    if (flush() != 0) return -1;
    const size_t n = id_to_label_.size();
    const size_t w0 = 2 * M_ + 1, wu = M_ + 1;
    std::vector<uint32_t> levels(n), l0(n * w0), upper;
    size_t recs = 0;
    long entry = -1, maxl = -1;
    if (n) {
        if (vsgpu_hnsw_export(graph_, levels.data(), l0.data(), nullptr, 0, &recs) != VSGPU_OK) return -1;
        upper.resize(recs * wu);
        if (recs && vsgpu_hnsw_export(graph_, nullptr, nullptr, upper.data(), recs, &recs) != VSGPU_OK) return -1;
        vsgpu_hnsw_entry(graph_, &entry, &maxl);
    }
    std::vector<uint8_t> rows(n * stored_size_);
    if (n && vsgpu_store_read(store_, 0, n, rows.data(), stored_size_, nullptr) != VSGPU_OK) return -1;
    // per (node, level): record offset, and who points at it without being pointed back at
    std::vector<const uint32_t *> rec_of;
    std::vector<size_t> first_rec(n + 1, 0);
    for (size_t i = 0; i < n; i++) first_rec[i + 1] = first_rec[i] + levels[i] + 1;
    rec_of.resize(first_rec[n]);
    size_t u = 0;
    for (size_t i = 0; i < n; i++) {
        rec_of[first_rec[i]] = l0.data() + i * w0;
        for (uint32_t l = 1; l <= levels[i]; l++) rec_of[first_rec[i] + l] = upper.data() + (u++) * wu;
    }
    std::vector<std::vector<idType>> incoming(first_rec[n]);
    for (size_t i = 0; i < n; i++)
        for (uint32_t l = 0; l <= levels[i]; l++) {
            const uint32_t *rec = rec_of[first_rec[i] + l];
            for (uint32_t j = 0; j < rec[0]; j++) {
                const uint32_t v = rec[1 + j];
                if (v >= n || levels[v] < l) continue;
                const uint32_t *back = rec_of[first_rec[v] + l];
                bool mutual = false;
                for (uint32_t t = 0; t < back[0]; t++)
                    if (back[1 + t] == (uint32_t)i) {
                        mutual = true;
                        break;
                    }
                if (!mutual) incoming[first_rec[v] + l].push_back((idType)i);
            }
        }
    std::ofstream out(path, std::ios::binary);
    if (!out.is_open()) return -1;
    put<int>(out, 4); // EncodingVersion::V4
    put<int>(out, VecSimAlgo_HNSWLIB);
    put<size_t>(out, dim_);
    put<int>(out, type_);
    put<int>(out, metric_);
    put<size_t>(out, block_size_);
    put<bool>(out, multi_);
    put<size_t>(out, (n + block_size_ - 1) / block_size_ * block_size_); // maxElements: whole blocks
    put<size_t>(out, M_);
    put<size_t>(out, 2 * M_);
    put<size_t>(out, efc_);
    put<size_t>(out, ef_);
    put<double>(out, epsilon_);
    put<double>(out, mult_);
    put<size_t>(out, n);
    put<size_t>(out, num_deleted_);
    put<size_t>(out, maxl < 0 ? (size_t)-1 : (size_t)maxl); // HNSW_INVALID_LEVEL
    put<idType>(out, entry < 0 ? (idType)-1 : (idType)entry);
    for (size_t i = 0; i < n; i++) {
        put<size_t>(out, id_to_label_[i]);
        put<uint8_t>(out, id_deleted_[i] ? DELETE_MARK : 0);
    }
    out.write((const char *)rows.data(), (std::streamsize)rows.size());
    for (size_t b = 0; b * block_size_ < n; b++) {
        const size_t first = b * block_size_, len = std::min(block_size_, n - first);
        put<unsigned int>(out, (unsigned int)len);
        for (size_t i = first; i < first + len; i++) {
            put<size_t>(out, levels[i]);
            for (uint32_t l = 0; l <= levels[i]; l++) {
                const uint32_t *rec = rec_of[first_rec[i] + l];
                put<uint16_t>(out, (uint16_t)rec[0]);
                out.write((const char *)(rec + 1), (std::streamsize)(rec[0] * sizeof(idType)));
                const auto &inc = incoming[first_rec[i] + l];
                put<unsigned int>(out, (unsigned int)inc.size());
                out.write((const char *)inc.data(), (std::streamsize)(inc.size() * sizeof(idType)));
            }
        }
    }
    return out.good() ? 0 : -1;
}

} // namespace vsb
