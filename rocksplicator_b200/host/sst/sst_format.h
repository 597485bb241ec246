// sst_format.h — RocksDB block-based SST files on the host: read (ingest) and write (spill), std only.
//
// NEXT-tier row of SURVEY.md §8(f) rank 2 ("durability + SST interchange"): the reference bulk-loads shards by
// ingesting external SST files (rocksdb_admin/admin_handler.cpp:1635-1853 addS3SstFilesToDB ->
// DB::IngestExternalFile) and pins cross-version compatibility with a golden file
// (rocksdb_admin/tests/old_sst_data.sst, sst_binary.cpp:43-78).  This header reads such files (block-based table,
// footer versions 0-2, Snappy or uncompressed blocks, CRC32C block checksums, prefix-compressed keys) and writes
// files RocksDB's own IngestExternalFile accepts (what SstFileWriter emits: sequence 0 keys, the two
// rocksdb.external_sst_file.* properties).  The format is RocksDB's published table format
// (table/format.h, table/block.h, table/block_based_table_builder.cc of the pinned 5.x line); parity is pinned by
// tests/test_sst_cpu.py against the reference's golden file and against files written / ingested by the
// reference's own librocksdb.so.5.4.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace sst {

struct Entry {
  std::string user_key;
  uint64_t seq = 0;
  uint8_t type = 1;  // RocksDB ValueType: 1 Put, 0 Delete, 2 Merge, 7 SingleDelete
  std::string value;
};
struct Props {
  std::map<std::string, std::string> raw;  // every property, value bytes as stored
  uint64_t num_entries = 0;
  uint32_t external_version = 0;  // rocksdb.external_sst_file.version (0 = not an external file)
  uint64_t global_seqno = 0;      // rocksdb.external_sst_file.global_seqno
  uint32_t format_version = 0;
};

static const uint64_t kBlockBasedMagic = 0x88e241b785f4cff7ull;
static const uint64_t kLegacyBlockBasedMagic = 0xdb4775248b80fb57ull;

// ---- coding -----------------------------------------------------------------------------------------
inline bool GetVarint64(const uint8_t** p, const uint8_t* lim, uint64_t* v) {
  uint64_t r = 0;
  for (uint32_t shift = 0; shift <= 63 && *p < lim; shift += 7) {
    const uint64_t b = *(*p)++;
    if (b & 128) r |= (b & 127) << shift;
    else { *v = r | (b << shift); return true; }
  }
  return false;
}
inline bool GetVarint32(const uint8_t** p, const uint8_t* lim, uint32_t* v) {
  uint64_t t;
  if (!GetVarint64(p, lim, &t) || t > 0xffffffffull) return false;
  *v = (uint32_t)t;
  return true;
}
inline void PutVarint64(std::string* s, uint64_t v) {
  while (v >= 128) { s->push_back((char)((v & 127) | 128)); v >>= 7; }
  s->push_back((char)v);
}
inline void PutFixed32(std::string* s, uint32_t v) { s->append((const char*)&v, 4); }
inline void PutFixed64(std::string* s, uint64_t v) { s->append((const char*)&v, 8); }
inline uint32_t DecodeFixed32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t DecodeFixed64(const void* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// ---- CRC32C (Castagnoli), masked as RocksDB stores it ------------------------------------------------------
inline uint32_t Crc32c(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
      table[i] = c;
    }
    init = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return ~crc;
}
inline uint32_t MaskCrc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// ---- Snappy decompression (format description: framing-free raw Snappy) -----------------------------------
inline bool SnappyUncompress(const uint8_t* p, size_t n, std::string* out) {
  const uint8_t* lim = p + n;
  uint64_t ulen;
  // (an untrusted length: a block never inflates beyond a few hundred times its stored size, and blocks are KBs)
  if (!GetVarint64(&p, lim, &ulen) || ulen > (64ull << 20) || ulen > 256ull * n + 4096) return false;
  out->clear();
  out->reserve((size_t)ulen);
  while (p < lim) {
    const uint8_t tag = *p++;
    const uint32_t kind = tag & 3;
    if (kind == 0) {  // literal
      uint32_t len = (tag >> 2) + 1;
      if (len > 60) {
        const uint32_t nb = len - 60;
        if ((size_t)(lim - p) < nb) return false;
        len = 0;
        for (uint32_t i = 0; i < nb; i++) len |= (uint32_t)p[i] << (8 * i);
        len += 1;
        p += nb;
      }
      if ((size_t)(lim - p) < len) return false;
      out->append((const char*)p, len);
      p += len;
    } else {
      uint32_t len, off;
      if (kind == 1) {
        if (p >= lim) return false;
        len = 4 + ((tag >> 2) & 7);
        off = ((uint32_t)(tag >> 5) << 8) | *p++;
      } else if (kind == 2) {
        if (lim - p < 2) return false;
        len = (tag >> 2) + 1;
        off = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
        p += 2;
      } else {
        if (lim - p < 4) return false;
        len = (tag >> 2) + 1;
        off = DecodeFixed32(p);
        p += 4;
      }
      if (off == 0 || off > out->size()) return false;
      const size_t start = out->size() - off;
      for (uint32_t i = 0; i < len; i++) out->push_back((*out)[start + i]);  // may overlap: byte by byte
    }
  }
  return out->size() == ulen;
}

// ---- blocks -------------------------------------------------------------------------------------------
struct BlockHandle {
  uint64_t offset = 0, size = 0;
  bool Decode(const uint8_t** p, const uint8_t* lim) { return GetVarint64(p, lim, &offset) && GetVarint64(p, lim, &size); }
  void Encode(std::string* s) const { PutVarint64(s, offset); PutVarint64(s, size); }
};

// read the block at `h` (verifying its checksum when present) into its uncompressed contents
inline bool ReadBlock(const std::string& file, const BlockHandle& h, bool verify, std::string* contents, std::string* err) {
  // (the handle's varint64s come from the file: no sums that can wrap)
  if (h.offset > file.size() || h.size > file.size() - h.offset || file.size() - h.offset - h.size < 5) { *err = "block handle out of range"; return false; }
  const uint8_t* p = (const uint8_t*)file.data() + h.offset;
  const uint8_t type = p[h.size];
  if (verify) {
    const uint32_t stored = DecodeFixed32(p + h.size + 1);
    const uint32_t actual = MaskCrc(Crc32c(0, p, h.size + 1));
    if (stored != actual) { *err = "block checksum mismatch"; return false; }
  }
  if (type == 0) { contents->assign((const char*)p, h.size); return true; }
  if (type == 1) {
    if (!SnappyUncompress(p, h.size, contents)) { *err = "corrupted snappy block"; return false; }
    return true;
  }
  *err = "unsupported block compression type " + std::to_string(type);
  return false;
}

// walk the key/value pairs of an uncompressed block
template <class F>
inline bool ForEachInBlock(const std::string& block, F f, std::string* err) {
  if (block.size() < 4) { *err = "block too small"; return false; }
  const uint32_t num_restarts = DecodeFixed32(block.data() + block.size() - 4);
  if ((uint64_t)num_restarts * 4 + 4 > block.size()) { *err = "bad restart count"; return false; }
  const uint8_t* p = (const uint8_t*)block.data();
  const uint8_t* lim = p + block.size() - 4 - (size_t)num_restarts * 4;
  std::string key;
  while (p < lim) {
    uint32_t shared, non_shared, vlen;
    if (!GetVarint32(&p, lim, &shared) || !GetVarint32(&p, lim, &non_shared) || !GetVarint32(&p, lim, &vlen) ||
        shared > key.size() || (size_t)(lim - p) < (size_t)non_shared + vlen) { *err = "bad block entry"; return false; }
    key.resize(shared);
    key.append((const char*)p, non_shared);
    p += non_shared;
    if (!f(key, std::string((const char*)p, vlen))) return false;
    p += vlen;
  }
  return true;
}

class BlockBuilder {
 public:
  explicit BlockBuilder(int restart_interval) : interval_(restart_interval) { restarts_.push_back(0); }
  void Add(const std::string& key, const std::string& value) {
    size_t shared = 0;
    if (counter_ < interval_) {
      const size_t m = std::min(last_key_.size(), key.size());
      while (shared < m && last_key_[shared] == key[shared]) shared++;
    } else {
      restarts_.push_back((uint32_t)buf_.size());
      counter_ = 0;
    }
    PutVarint64(&buf_, shared);
    PutVarint64(&buf_, key.size() - shared);
    PutVarint64(&buf_, value.size());
    buf_.append(key.data() + shared, key.size() - shared);
    buf_.append(value);
    last_key_ = key;
    counter_++;
    n_++;
  }
  std::string Finish() {
    std::string out = buf_;
    for (uint32_t r : restarts_) PutFixed32(&out, r);
    PutFixed32(&out, (uint32_t)restarts_.size());
    return out;
  }
  size_t EstimatedSize() const { return buf_.size() + restarts_.size() * 4 + 4; }
  bool empty() const { return n_ == 0; }

 private:
  int interval_, counter_ = 0;
  size_t n_ = 0;
  std::string buf_, last_key_;
  std::vector<uint32_t> restarts_;
};

// ---- reader ---------------------------------------------------------------------------------------------
inline bool ReadSst(const std::string& file, std::vector<Entry>* out, Props* props, std::string* err, bool verify = true) {
  out->clear();
  if (file.size() < 48) { *err = "file too short to be an sstable"; return false; }
  const uint64_t magic = DecodeFixed64(file.data() + file.size() - 8);
  const uint8_t* p;
  const uint8_t* lim;
  uint32_t version = 0;
  if (magic == kBlockBasedMagic) {
    if (file.size() < 53) { *err = "file too short"; return false; }
    const uint8_t* footer = (const uint8_t*)file.data() + file.size() - 53;
    version = DecodeFixed32(footer + 41);
    p = footer + 1;  // footer[0] = checksum type (1 = CRC32C)
    lim = footer + 41;
    if (footer[0] != 1) verify = false;
  } else if (magic == kLegacyBlockBasedMagic) {
    p = (const uint8_t*)file.data() + file.size() - 48;
    lim = p + 40;
  } else {
    *err = "not a block-based sstable (bad magic number)";
    return false;
  }
  BlockHandle metaindex, index;
  if (!metaindex.Decode(&p, lim) || !index.Decode(&p, lim)) { *err = "bad footer handles"; return false; }
  if (props) props->format_version = version;
  // properties, through the metaindex
  std::string blk;
  if (props) {
    if (!ReadBlock(file, metaindex, verify, &blk, err)) return false;
    BlockHandle ph;
    bool have = false;
    if (!ForEachInBlock(blk, [&](const std::string& k, const std::string& v) {
          if (k == "rocksdb.properties") {
            const uint8_t* q = (const uint8_t*)v.data();
            have = ph.Decode(&q, q + v.size());
          }
          return true;
        }, err)) return false;
    if (have) {
      std::string pb;
      if (!ReadBlock(file, ph, verify, &pb, err)) return false;
      if (!ForEachInBlock(pb, [&](const std::string& k, const std::string& v) { props->raw[k] = v; return true; }, err)) return false;
      auto num = props->raw.find("rocksdb.num.entries");
      if (num != props->raw.end()) {
        const uint8_t* q = (const uint8_t*)num->second.data();
        GetVarint64(&q, q + num->second.size(), &props->num_entries);
      }
      auto ev = props->raw.find("rocksdb.external_sst_file.version");
      if (ev != props->raw.end() && ev->second.size() >= 4) props->external_version = DecodeFixed32(ev->second.data());
      auto gs = props->raw.find("rocksdb.external_sst_file.global_seqno");
      if (gs != props->raw.end() && gs->second.size() >= 8) props->global_seqno = DecodeFixed64(gs->second.data());
    }
  }
  // data blocks in index order
  std::string index_blk;
  if (!ReadBlock(file, index, verify, &index_blk, err)) return false;
  std::vector<BlockHandle> blocks;
  if (!ForEachInBlock(index_blk, [&](const std::string&, const std::string& v) {
        BlockHandle h;
        const uint8_t* q = (const uint8_t*)v.data();
        if (!h.Decode(&q, q + v.size())) return false;
        blocks.push_back(h);
        return true;
      }, err)) { if (err->empty()) *err = "bad index entry"; return false; }
  for (const BlockHandle& h : blocks) {
    std::string data;
    if (!ReadBlock(file, h, verify, &data, err)) return false;
    if (!ForEachInBlock(data, [&](const std::string& ikey, const std::string& v) {
          if (ikey.size() < 8) return false;
          Entry e;
          e.user_key.assign(ikey.data(), ikey.size() - 8);
          const uint64_t st = DecodeFixed64(ikey.data() + ikey.size() - 8);
          e.seq = st >> 8;
          e.type = (uint8_t)(st & 0xff);
          e.value = v;
          out->push_back(std::move(e));
          return true;
        }, err)) { if (err->empty()) *err = "bad data entry"; return false; }
  }
  // an ingested external file carries one global sequence number for all its keys
  if (props && props->external_version == 2 && props->global_seqno)
    for (Entry& e : *out) e.seq = props->global_seqno;
  return true;
}

// ---- writer: what SstFileWriter emits (keys in strictly increasing order, sequence 0, Puts) ---------------------
inline void AppendBlock(std::string* file, const std::string& contents, BlockHandle* h) {
  h->offset = file->size();
  h->size = contents.size();
  file->append(contents);
  file->push_back(0);  // kNoCompression
  const uint32_t crc = MaskCrc(Crc32c(0, (const uint8_t*)file->data() + h->offset, contents.size() + 1));
  PutFixed32(file, crc);
}

inline bool WriteSst(const std::vector<std::pair<std::string, std::string>>& sorted_kv, std::string* file, std::string* err,
                     size_t block_size = 4096) {
  file->clear();
  for (size_t i = 1; i < sorted_kv.size(); i++)
    if (!(sorted_kv[i - 1].first < sorted_kv[i].first)) { *err = "Keys must be added in order"; return false; }
  BlockBuilder index(1);
  uint64_t raw_key = 0, raw_val = 0, n_blocks = 0;
  BlockBuilder data(16);
  std::string last_ikey;
  auto flush = [&] {
    if (data.empty()) return;
    BlockHandle h;
    AppendBlock(file, data.Finish(), &h);
    std::string hv;
    h.Encode(&hv);
    index.Add(last_ikey, hv);  // separator = the block's last key
    n_blocks++;
    data = BlockBuilder(16);
  };
  for (const auto& kv : sorted_kv) {
    std::string ikey = kv.first;
    PutFixed64(&ikey, (0ull << 8) | 1ull);  // sequence 0, kTypeValue
    data.Add(ikey, kv.second);
    last_ikey = ikey;
    raw_key += ikey.size();
    raw_val += kv.second.size();
    if (data.EstimatedSize() >= block_size) flush();
  }
  flush();
  const uint64_t data_size = file->size();
  const std::string index_contents = index.Finish();
  // properties (sorted by name; numeric values are varint64, the two external-file properties fixed-width)
  std::map<std::string, std::string> pm;
  auto num = [&](const char* name, uint64_t v) { std::string s; PutVarint64(&s, v); pm[name] = s; };
  num("rocksdb.data.size", data_size);
  num("rocksdb.index.size", index_contents.size() + 5);
  num("rocksdb.raw.key.size", raw_key);
  num("rocksdb.raw.value.size", raw_val);
  num("rocksdb.num.data.blocks", n_blocks);
  num("rocksdb.num.entries", sorted_kv.size());
  num("rocksdb.filter.size", 0);
  num("rocksdb.format.version", 0);
  num("rocksdb.fixed.key.length", 0);
  pm["rocksdb.comparator"] = "leveldb.BytewiseComparator";
  pm["rocksdb.compression"] = "NoCompression";
  pm["rocksdb.filter.policy"] = "";
  pm["rocksdb.merge.operator"] = "nullptr";
  pm["rocksdb.prefix.extractor.name"] = "nullptr";
  pm["rocksdb.property.collectors"] = "[SstFileWriterCollector]";
  { std::string v; PutFixed32(&v, 2); pm["rocksdb.external_sst_file.version"] = v; }
  { std::string v; PutFixed64(&v, 0); pm["rocksdb.external_sst_file.global_seqno"] = v; }
  BlockBuilder pb(1);
  for (const auto& kv : pm) pb.Add(kv.first, kv.second);
  BlockHandle ph, mh, ih;
  AppendBlock(file, pb.Finish(), &ph);
  BlockBuilder mb(1);
  { std::string hv; ph.Encode(&hv); mb.Add("rocksdb.properties", hv); }
  AppendBlock(file, mb.Finish(), &mh);
  AppendBlock(file, index_contents, &ih);
  // footer, format version 2: [checksum type][metaindex handle][index handle][pad to 41][version][magic]
  std::string footer;
  footer.push_back(1);  // kCRC32c
  mh.Encode(&footer);
  ih.Encode(&footer);
  footer.resize(41, '\0');
  PutFixed32(&footer, 2);
  PutFixed64(&footer, kBlockBasedMagic);
  file->append(footer);
  return true;
}

}  // namespace sst
