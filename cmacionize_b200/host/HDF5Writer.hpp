/*
 * HDF5Writer.hpp — a self-contained writer for the subset of the HDF5 file format that the reference's
 * snapshot files use, so that the Gadget-style snapshots (GadgetDensityGridWriter.cpp:122-208) can be
 * written without an HDF5 library (the image has none; the reference links libhdf5 through HDF5Tools.hpp).
 *
 * What is written is the classic ("1.6 compatible") on-disk layout that every HDF5 release reads:
 *   superblock version 0 (8-byte offsets and lengths, group leaf K = 4, internal K = 16),
 *   groups as symbol tables: version-1 object header + B-tree (v1, node type 0) + local heap + symbol nodes,
 *   attributes as version-1 attribute messages in the object header (scalar or 1-D; IEEE double,
 *   32/64-bit integers, null-terminated fixed-length strings = H5T_C_S1 with size strlen + 1),
 *   datasets with a version-1 simple dataspace, a version-2 fill-value message and a version-3 CONTIGUOUS
 *   layout (the reference writes chunked datasets; the layout is invisible to readers).
 * The byte layouts of the messages were checked against files written by the library itself (the reference's
 * test/test.hdf5 and test/taskbased.hdf5) with tests/h5mini.py, which is also what reads the files back in
 * tests/test_hdf5_writer.py.
 *
 * Use: build a tree (HDF5File::root().create_group(..).write_attribute(..) / .create_dataset(..)), then
 * write(filename).  Dataset memory is borrowed until write() returns.
 */
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <string>
#include <vector>

#include "Error.hpp"

namespace cmi {
namespace hdf5 {

enum class Type { F64, I32, U32, U64, STRING };

inline size_t type_size(Type t, size_t string_size = 0) {
  switch (t) {
  case Type::F64: return 8;
  case Type::I32: return 4;
  case Type::U32: return 4;
  case Type::U64: return 8;
  case Type::STRING: return string_size;
  }
  return 0;
}

class Buffer {
public:
  std::vector<uint8_t> bytes;
  void u8(uint8_t v) { bytes.push_back(v); }
  void u16(uint16_t v) { raw(&v, 2); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void raw(const void *p, size_t n) {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    bytes.insert(bytes.end(), b, b + n);
  }
  void zeros(size_t n) { bytes.insert(bytes.end(), n, 0); }
  void pad8() { zeros((8 - bytes.size() % 8) % 8); }
  size_t size() const { return bytes.size(); }
};

inline size_t pad8(size_t n) { return (n + 7) & ~size_t(7); }

/* datatype message body */
inline void put_datatype(Buffer &b, Type t, size_t string_size) {
  switch (t) {
  case Type::F64: /* class 1, version 1; little endian, implied mantissa msb, sign bit 63 */
    b.u8(0x11); b.u8(0x20); b.u8(0x3f); b.u8(0x00); b.u32(8);
    b.u16(0); b.u16(64); b.u8(52); b.u8(11); b.u8(0); b.u8(52); b.u32(1023);
    break;
  case Type::I32:
    b.u8(0x10); b.u8(0x08); b.u8(0); b.u8(0); b.u32(4); b.u16(0); b.u16(32);
    break;
  case Type::U32:
    b.u8(0x10); b.u8(0x00); b.u8(0); b.u8(0); b.u32(4); b.u16(0); b.u16(32);
    break;
  case Type::U64:
    b.u8(0x10); b.u8(0x00); b.u8(0); b.u8(0); b.u32(8); b.u16(0); b.u16(64);
    break;
  case Type::STRING: /* class 3: null terminated, ASCII */
    b.u8(0x13); b.u8(0x00); b.u8(0); b.u8(0); b.u32((uint32_t)string_size);
    break;
  }
}

/* dataspace message body, version 1; rank 0 = scalar; simple spaces carry their maximum dimensions */
inline void put_dataspace(Buffer &b, const std::vector<uint64_t> &dims) {
  b.u8(1); b.u8((uint8_t)dims.size()); b.u8(dims.empty() ? 0 : 1); b.zeros(5);
  for (uint64_t d : dims) b.u64(d);
  for (uint64_t d : dims) b.u64(d);
}

struct Attribute {
  std::string name;
  Type type;
  size_t string_size = 0;
  std::vector<uint64_t> dims; /* empty = scalar */
  std::vector<uint8_t> data;

  /* the whole attribute message (version 1) */
  void put(Buffer &out) const {
    Buffer dt, ds;
    put_datatype(dt, type, string_size);
    put_dataspace(ds, dims);
    out.u8(1); out.u8(0);
    out.u16((uint16_t)(name.size() + 1));
    out.u16((uint16_t)dt.size());
    out.u16((uint16_t)ds.size());
    out.raw(name.c_str(), name.size() + 1); out.pad8();
    out.raw(dt.bytes.data(), dt.size()); out.pad8();
    out.raw(ds.bytes.data(), ds.size()); out.pad8();
    out.raw(data.data(), data.size());
  }
};

struct Dataset {
  std::string name;
  Type type;
  std::vector<uint64_t> dims;
  const void *data;
  uint64_t header_address = 0, data_address = 0;
  uint64_t nbytes() const {
    uint64_t n = type_size(type);
    for (uint64_t d : dims) n *= d;
    return n;
  }
};

class Group {
public:
  explicit Group(std::string name) : name_(std::move(name)) {}

  Group &create_group(const std::string &name) {
    check_new_link(name);
    groups_.emplace_back(new Group(name));
    return *groups_.back();
  }
  /* memory at `data` must stay valid until HDF5File::write returns */
  void create_dataset(const std::string &name, Type type, const std::vector<uint64_t> &dims, const void *data) {
    check_new_link(name);
    if (type == Type::STRING) cmi_error("String datasets are not provided!");
    datasets_.push_back(Dataset{name, type, dims, data});
  }

  void write_attribute(const std::string &name, double v) { scalar(name, Type::F64, &v, 8); }
  void write_attribute(const std::string &name, int32_t v) { scalar(name, Type::I32, &v, 4); }
  void write_attribute(const std::string &name, uint32_t v) { scalar(name, Type::U32, &v, 4); }
  void write_attribute(const std::string &name, uint64_t v) { scalar(name, Type::U64, &v, 8); }
  void write_attribute(const std::string &name, const std::string &v) {
    Attribute a;
    a.name = name; a.type = Type::STRING; a.string_size = v.size() + 1;
    a.data.assign(v.c_str(), v.c_str() + v.size() + 1);
    add(std::move(a));
  }
  void write_attribute(const std::string &name, const char *v) { write_attribute(name, std::string(v)); }
  void write_attribute(const std::string &name, const std::vector<double> &v) { vector(name, Type::F64, v.data(), v.size(), 8); }
  void write_attribute(const std::string &name, const std::vector<uint32_t> &v) { vector(name, Type::U32, v.data(), v.size(), 4); }
  void write_attribute(const std::string &name, const std::vector<int32_t> &v) { vector(name, Type::I32, v.data(), v.size(), 4); }
  void write_attribute(const std::string &name, const std::array<double, 3> &v) { vector(name, Type::F64, v.data(), 3, 8); }

private:
  friend class HDF5File;
  struct Link {
    std::string name;
    Group *group;
    Dataset *dataset;
    uint64_t heap_offset;
  };

  void check_new_link(const std::string &name) const {
    if (name.empty() || name.find('/') != std::string::npos) cmi_error("Invalid HDF5 link name \"%s\"!", name.c_str());
    for (const auto &g : groups_)
      if (g->name_ == name) cmi_error("HDF5 link \"%s\" already exists!", name.c_str());
    for (const auto &d : datasets_)
      if (d.name == name) cmi_error("HDF5 link \"%s\" already exists!", name.c_str());
  }
  void add(Attribute &&a) {
    for (const auto &o : attributes_)
      if (o.name == a.name) cmi_error("HDF5 attribute \"%s\" already exists!", a.name.c_str());
    Buffer probe;
    a.put(probe);
    if (probe.size() > 65000) cmi_error("HDF5 attribute \"%s\" is too large for an object header message!", a.name.c_str());
    attributes_.push_back(std::move(a));
  }
  void scalar(const std::string &name, Type t, const void *p, size_t n) {
    Attribute a;
    a.name = name; a.type = t;
    a.data.assign((const uint8_t *)p, (const uint8_t *)p + n);
    add(std::move(a));
  }
  void vector(const std::string &name, Type t, const void *p, size_t count, size_t elem) {
    Attribute a;
    a.name = name; a.type = t; a.dims = {count};
    a.data.assign((const uint8_t *)p, (const uint8_t *)p + count * elem);
    add(std::move(a));
  }

  std::string name_;
  std::vector<Attribute> attributes_;
  std::vector<std::unique_ptr<Group>> groups_;
  std::vector<Dataset> datasets_;
  /* layout, filled by HDF5File */
  std::vector<Link> links_; /* sorted by name */
  std::vector<uint8_t> heap_data_;
  uint64_t heap_free_offset_ = 0;
  uint64_t header_address_ = 0, btree_address_ = 0, heap_address_ = 0, heap_data_address_ = 0;
  std::vector<uint64_t> snod_address_;
};

class HDF5File {
public:
  HDF5File() : root_("") {}
  Group &root() { return root_; }

  void write(const std::string &filename) {
    uint64_t eof = plan();
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) cmi_error("Unable to open file \"%s\" for writing!", filename.c_str());
    file_ = f;
    position_ = 0;
    mod_time_ = (uint32_t)time(nullptr);
    Buffer sb;
    static const uint8_t signature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    sb.raw(signature, 8);
    sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); /* versions: superblock, free space, root entry, -, shared header */
    sb.u8(8); sb.u8(8); sb.u8(0);                     /* size of offsets, size of lengths */
    sb.u16(LEAF_K); sb.u16(INTERNAL_K);
    sb.u32(0);                                        /* consistency flags */
    sb.u64(0); sb.u64(UNDEFINED); sb.u64(eof); sb.u64(UNDEFINED); /* base, free-space info, end of file, driver info */
    put_symbol_entry(sb, 0, root_);                   /* root group symbol table entry */
    emit(sb, 0);
    write_group(root_);
    if (position_ != eof) cmi_error("HDF5 writer: wrote %llu bytes, planned %llu!", (unsigned long long)position_, (unsigned long long)eof);
    if (fclose(f) != 0) cmi_error("Error while closing file \"%s\"!", filename.c_str());
    file_ = nullptr;
  }

private:
  static constexpr uint16_t LEAF_K = 4, INTERNAL_K = 16;
  static constexpr uint64_t UNDEFINED = ~uint64_t(0);
  static constexpr uint64_t SUPERBLOCK_SIZE = 96;
  static constexpr uint64_t BTREE_NODE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8;
  static constexpr uint64_t SNOD_SIZE = 8 + 2 * LEAF_K * 40;
  static constexpr uint64_t HEAP_HEADER_SIZE = 32;

  /* ---- sizes ---- */
  static uint64_t message_size(size_t body) { return 8 + pad8(body); }
  static uint64_t group_header_size(const Group &g) {
    uint64_t n = 16 + message_size(16);
    for (const Attribute &a : g.attributes_) {
      Buffer b;
      a.put(b);
      n += message_size(b.size());
    }
    return n;
  }
  static uint64_t dataset_header_size(const Dataset &d) {
    Buffer ds, dt;
    put_dataspace(ds, d.dims);
    put_datatype(dt, d.type, 0);
    return 16 + message_size(ds.size()) + message_size(dt.size()) + message_size(8) + message_size(18) + message_size(8);
  }

  /* ---- pass 1: addresses, in the order pass 2 emits ---- */
  uint64_t plan() {
    uint64_t at = SUPERBLOCK_SIZE;
    plan_group(root_, at);
    return at;
  }
  void plan_group(Group &g, uint64_t &at) {
    g.links_.clear();
    for (auto &c : g.groups_) g.links_.push_back({c->name_, c.get(), nullptr, 0});
    for (auto &d : g.datasets_) g.links_.push_back({d.name, nullptr, &d, 0});
    std::sort(g.links_.begin(), g.links_.end(), [](const Group::Link &a, const Group::Link &b) { return strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    if (g.links_.size() > size_t(2 * INTERNAL_K) * 2 * LEAF_K) cmi_error("Too many links in one HDF5 group!");
    /* local heap: the empty name at offset 0, then the link names, each padded to 8 bytes, then one free block */
    g.heap_data_.assign(8, 0);
    for (auto &l : g.links_) {
      l.heap_offset = g.heap_data_.size();
      g.heap_data_.insert(g.heap_data_.end(), l.name.begin(), l.name.end());
      g.heap_data_.push_back(0);
      g.heap_data_.resize(pad8(g.heap_data_.size()), 0);
    }
    g.heap_free_offset_ = g.heap_data_.size();
    const uint64_t free_size = 32;
    Buffer fb;
    fb.u64(1); /* no next free block */
    fb.u64(free_size);
    fb.zeros(free_size - 16);
    g.heap_data_.insert(g.heap_data_.end(), fb.bytes.begin(), fb.bytes.end());

    g.header_address_ = at; at += group_header_size(g);
    g.btree_address_ = at; at += BTREE_NODE_SIZE;
    g.heap_address_ = at; at += HEAP_HEADER_SIZE;
    g.heap_data_address_ = at; at += g.heap_data_.size();
    const size_t nsnod = (g.links_.size() + 2 * LEAF_K - 1) / (2 * LEAF_K);
    g.snod_address_.assign(nsnod, 0);
    for (size_t k = 0; k < nsnod; ++k) { g.snod_address_[k] = at; at += SNOD_SIZE; }
    for (auto &l : g.links_) {
      if (l.group) {
        plan_group(*l.group, at);
      } else {
        l.dataset->header_address = at; at += dataset_header_size(*l.dataset);
        l.dataset->data_address = at; at += pad8(l.dataset->nbytes());
      }
    }
  }

  /* ---- pass 2 ---- */
  void emit(const Buffer &b, uint64_t address) {
    if (address != position_) cmi_error("HDF5 writer: block planned at %llu is written at %llu!", (unsigned long long)address, (unsigned long long)position_);
    if (b.size() && fwrite(b.bytes.data(), 1, b.size(), file_) != b.size()) cmi_error("Error while writing HDF5 file!");
    position_ += b.size();
  }
  static void put_message_header(Buffer &b, uint16_t type, size_t body, uint8_t flags) {
    b.u16(type); b.u16((uint16_t)pad8(body)); b.u8(flags); b.zeros(3);
  }
  static void put_symbol_entry(Buffer &b, uint64_t name_offset, const Group &g) {
    b.u64(name_offset); b.u64(g.header_address_); b.u32(1); b.u32(0); /* cached: B-tree and heap addresses */
    b.u64(g.btree_address_); b.u64(g.heap_address_);
  }
  void write_group(Group &g) {
    /* object header */
    Buffer h;
    h.u8(1); h.u8(0); h.u16((uint16_t)(1 + g.attributes_.size())); h.u32(1);
    h.u32((uint32_t)(group_header_size(g) - 16)); h.u32(0);
    put_message_header(h, 0x0011, 16, 0);
    h.u64(g.btree_address_); h.u64(g.heap_address_);
    for (const Attribute &a : g.attributes_) {
      Buffer body;
      a.put(body);
      put_message_header(h, 0x000C, body.size(), 0);
      h.raw(body.bytes.data(), body.size()); h.pad8();
    }
    emit(h, g.header_address_);
    /* B-tree: one leaf-level node whose children are the symbol nodes */
    Buffer t;
    t.raw("TREE", 4); t.u8(0); t.u8(0); t.u16((uint16_t)g.snod_address_.size());
    t.u64(UNDEFINED); t.u64(UNDEFINED);
    const size_t per = 2 * LEAF_K;
    t.u64(0); /* key 0: the empty name */
    for (size_t k = 0; k < g.snod_address_.size(); ++k) {
      t.u64(g.snod_address_[k]);
      const size_t last = std::min(g.links_.size(), (k + 1) * per) - 1;
      t.u64(g.links_[last].heap_offset); /* key k + 1: largest name in child k */
    }
    t.zeros(BTREE_NODE_SIZE - t.size());
    emit(t, g.btree_address_);
    /* local heap */
    Buffer hp;
    hp.raw("HEAP", 4); hp.u8(0); hp.zeros(3);
    hp.u64(g.heap_data_.size()); hp.u64(g.heap_free_offset_); hp.u64(g.heap_data_address_);
    emit(hp, g.heap_address_);
    Buffer hd;
    hd.raw(g.heap_data_.data(), g.heap_data_.size());
    emit(hd, g.heap_data_address_);
    /* symbol nodes */
    for (size_t k = 0; k < g.snod_address_.size(); ++k) {
      Buffer s;
      const size_t first = k * per, last = std::min(g.links_.size(), first + per);
      s.raw("SNOD", 4); s.u8(1); s.u8(0); s.u16((uint16_t)(last - first));
      for (size_t i = first; i < last; ++i) {
        const Group::Link &l = g.links_[i];
        if (l.group) {
          put_symbol_entry(s, l.heap_offset, *l.group);
        } else {
          s.u64(l.heap_offset); s.u64(l.dataset->header_address); s.u32(0); s.u32(0); s.zeros(16);
        }
      }
      s.zeros(SNOD_SIZE - s.size());
      emit(s, g.snod_address_[k]);
    }
    for (auto &l : g.links_) {
      if (l.group) write_group(*l.group);
      else write_dataset(*l.dataset);
    }
  }
  void write_dataset(const Dataset &d) {
    Buffer ds, dt;
    put_dataspace(ds, d.dims);
    put_datatype(dt, d.type, 0);
    Buffer h;
    h.u8(1); h.u8(0); h.u16(5); h.u32(1); h.u32((uint32_t)(dataset_header_size(d) - 16)); h.u32(0);
    put_message_header(h, 0x0001, ds.size(), 0);
    h.raw(ds.bytes.data(), ds.size()); h.pad8();
    put_message_header(h, 0x0003, dt.size(), 1);
    h.raw(dt.bytes.data(), dt.size()); h.pad8();
    put_message_header(h, 0x0005, 8, 1); /* fill value, version 2: late allocation, written if set, default (undefined size 0) */
    h.u8(2); h.u8(2); h.u8(2); h.u8(1); h.u32(0);
    put_message_header(h, 0x0008, 18, 1); /* layout, version 3, contiguous */
    h.u8(3); h.u8(1); h.u64(d.data_address); h.u64(d.nbytes()); h.pad8();
    put_message_header(h, 0x0012, 8, 0); /* modification time */
    h.u8(1); h.zeros(3); h.u32(mod_time_);
    emit(h, d.header_address);
    if (position_ != d.data_address) cmi_error("HDF5 writer: data planned at %llu is written at %llu!", (unsigned long long)d.data_address, (unsigned long long)position_);
    const uint64_t n = d.nbytes();
    if (n && fwrite(d.data, 1, n, file_) != n) cmi_error("Error while writing HDF5 file!");
    static const uint8_t zero[8] = {0};
    const uint64_t pad = pad8(n) - n;
    if (pad && fwrite(zero, 1, pad, file_) != pad) cmi_error("Error while writing HDF5 file!");
    position_ += pad8(n);
  }

  Group root_;
  FILE *file_ = nullptr;
  uint64_t position_ = 0;
  uint32_t mod_time_ = 0;
};

} // namespace hdf5
} // namespace cmi
