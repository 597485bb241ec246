// sst_c_api.cpp — C entry points over sst_format.h (librsp_host.so), for bindings and tests.
#include <cstdlib>
#include <cstring>

#include "sst/sst_format.h"

extern "C" {

// Parse a block-based SST image.  *out receives malloc'ed packed records
// [u32 klen][u32 vlen][u64 seq][u8 type][key][value]...; props_out (optional, >= 3 u64): num_entries,
// external file version, global seqno.  Returns 0 on success.
int rsp_sst_read(const uint8_t* file, size_t len, uint8_t** out, size_t* out_len, size_t* n_entries, uint64_t* props_out,
                 char* err, size_t errcap) {
  std::vector<sst::Entry> entries;
  sst::Props props;
  std::string e;
  if (!sst::ReadSst(std::string((const char*)file, len), &entries, &props, &e)) {
    if (err && errcap) snprintf(err, errcap, "%s", e.c_str());
    return 2;
  }
  size_t total = 0;
  for (auto& x : entries) total += 17 + x.user_key.size() + x.value.size();
  uint8_t* buf = (uint8_t*)malloc(total ? total : 1);
  size_t at = 0;
  for (auto& x : entries) {
    const uint32_t kl = (uint32_t)x.user_key.size(), vl = (uint32_t)x.value.size();
    memcpy(buf + at, &kl, 4); memcpy(buf + at + 4, &vl, 4); memcpy(buf + at + 8, &x.seq, 8); buf[at + 16] = x.type;
    memcpy(buf + at + 17, x.user_key.data(), kl); memcpy(buf + at + 17 + kl, x.value.data(), vl);
    at += 17 + kl + vl;
  }
  *out = buf; *out_len = total; *n_entries = entries.size();
  if (props_out) { props_out[0] = props.num_entries; props_out[1] = props.external_version; props_out[2] = props.global_seqno; }
  return 0;
}

// Build an ingestible SST image from n sorted key/value pairs (keys concatenated with koff[n+1], values with voff[n+1]).
int rsp_sst_write(size_t n, const uint8_t* keys, const uint64_t* koff, const uint8_t* vals, const uint64_t* voff,
                  uint32_t block_size, uint8_t** out, size_t* out_len, char* err, size_t errcap) {
  std::vector<std::pair<std::string, std::string>> kv(n);
  for (size_t i = 0; i < n; i++) {
    kv[i].first.assign((const char*)keys + koff[i], (size_t)(koff[i + 1] - koff[i]));
    kv[i].second.assign((const char*)vals + voff[i], (size_t)(voff[i + 1] - voff[i]));
  }
  std::string file, e;
  if (!sst::WriteSst(kv, &file, &e, block_size ? block_size : 4096)) {
    if (err && errcap) snprintf(err, errcap, "%s", e.c_str());
    return 4;
  }
  *out = (uint8_t*)malloc(file.size());
  memcpy(*out, file.data(), file.size());
  *out_len = file.size();
  return 0;
}

void rsp_host_free(void* p) { free(p); }

}  // extern "C"
