/*
 * HDF5Reader.hpp — reads the snapshot files of the reference (written by libhdf5 through HDF5Tools.hpp) and
 * of host/HDF5Writer.hpp without an HDF5 library: the counterpart of HDF5Tools::open_file / open_group /
 * group_exists / get_attribute_names / read_attribute / read_dataset (HDF5Tools.hpp:176-1250) for the classic
 * on-disk format those files use:
 *   superblock version 0, version-1 object headers (with continuation blocks), symbol-table groups
 *   (B-tree v1 + symbol nodes + local heap), version-1 attribute messages, fixed-point / IEEE floating point /
 *   fixed-length string datatypes, contiguous, compact and chunked layouts (B-tree v1 of chunks), the deflate
 *   and shuffle filters (what `DensityGridWriter:compression` of the reference switches on; zlib).
 * Anything else (new-style groups, version-2 object headers, variable-length strings, ...) is refused with an
 * error that names it.  The file is memory mapped.  The same logic exists in Python as tests/h5mini.py.
 */
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "Error.hpp"

namespace cmi {
namespace hdf5 {

class HDF5Input {
public:
  explicit HDF5Input(const std::string &filename) : filename_(filename) {
    const int fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) cmi_error("Unable to open file \"%s\"!", filename.c_str());
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 96) {
      ::close(fd);
      cmi_error("File \"%s\" is not an HDF5 file!", filename.c_str());
    }
    size_ = (uint64_t)st.st_size;
    void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) cmi_error("Unable to map file \"%s\"!", filename.c_str());
    data_ = static_cast<const uint8_t *>(p);
    static const uint8_t signature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (memcmp(data_, signature, 8) != 0) fail("not an HDF5 file");
    if (data_[8] != 0) fail("superblock version " + std::to_string(data_[8]) + " (only the classic version 0 is read)");
    if (data_[13] != 8 || data_[14] != 8) fail("offsets / lengths that are not 8 bytes");
    if (u64(24) != 0) fail("a non-zero base address");
    root_ = parse_object(u64(56 + 8));
  }
  ~HDF5Input() {
    if (data_) munmap(const_cast<uint8_t *>(data_), size_);
  }
  HDF5Input(const HDF5Input &) = delete;
  HDF5Input &operator=(const HDF5Input &) = delete;

  /* HDF5Tools::group_exists: any link (group or dataset) under that path */
  bool exists(const std::string &path) const {
    uint64_t address;
    return find(path, address);
  }
  std::vector<std::string> get_attribute_names(const std::string &path) const {
    const Object o = open_object(path);
    std::vector<std::string> names;
    for (const Attribute &a : o.attributes) names.push_back(a.name);
    return names;
  }
  std::string read_string_attribute(const std::string &path, const std::string &name) const {
    const Attribute &a = attribute(path, name);
    if (a.type.cls != 3) fail("attribute \"" + name + "\" is not a string");
    check(a.data, a.type.size); /* a zero-sized dataspace checked nothing when the attribute was parsed */
    const char *s = reinterpret_cast<const char *>(data_ + a.data);
    return std::string(s, strnlen(s, a.type.size));
  }
  std::vector<double> read_double_attribute(const std::string &path, const std::string &name) const {
    const Attribute &a = attribute(path, name);
    std::vector<double> out(a.count);
    convert(a.type, data_ + a.data, a.count, out.data());
    return out;
  }
  /* HDF5Tools::read_dictionary (HDF5Tools.hpp:1228-1333): a 1-D dataset of {name, value} compounds (the runtime
   * parameter tables of FLASH) as a map; trailing blanks of the names are stripped */
  std::map<std::string, double> read_dictionary(const std::string &path) const {
    const Object o = open_object(path);
    if (!o.has_layout || !o.has_type || !o.has_space || o.type.cls != 6 || o.type.members.size() != 2 ||
        o.type.members[0].cls != 3 || (o.type.members[1].cls != 0 && o.type.members[1].cls != 1))
      fail("\"" + path + "\" is not a {name, value} table");
    if (o.layout_class != 1 || o.layout_address == UNDEFINED) fail("\"" + path + "\": only contiguous tables are read");
    const uint64_t n = element_count(o.dims);
    check(o.layout_address, mul(n, o.type.size));
    std::map<std::string, double> out;
    const Member &key = o.type.members[0], &val = o.type.members[1];
    if ((uint64_t)key.offset + key.size > o.type.size || (uint64_t)val.offset + val.size > o.type.size)
      fail("\"" + path + "\": a compound member outside its record");
    Datatype vt;
    vt.cls = val.cls; vt.size = val.size; vt.is_signed = val.is_signed;
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t *e = data_ + o.layout_address + i * o.type.size;
      const char *name = reinterpret_cast<const char *>(e + key.offset);
      std::string k(name, strnlen(name, key.size));
      while (!k.empty() && k.back() == ' ') k.pop_back();
      double v;
      convert(vt, e + val.offset, 1, &v);
      out[k] = v;
    }
    return out;
  }
  /* a dataset as doubles (whatever numeric type it has on disk), row major; dims gets its shape */
  std::vector<double> read_dataset(const std::string &path, std::vector<uint64_t> *dims = nullptr) const {
    const Object o = open_object(path);
    if (!o.has_layout || !o.has_type || !o.has_space) fail("\"" + path + "\" is not a dataset");
    if (dims) *dims = o.dims;
    const uint64_t n = element_count(o.dims);
    const uint64_t esize = o.type.size;
    const uint64_t nbytes = mul(n, esize);
    if (n > MAX_ELEMENTS) fail("dataset \"" + path + "\": a dataspace of more than 2^34 elements (corrupted?)");
    /* contiguous / compact data must be in the file before anything is allocated for it */
    if ((o.layout_class == 1 && o.layout_address != UNDEFINED) || o.layout_class == 0) {
      if (o.layout_size < nbytes) fail("dataset \"" + path + "\": layout smaller than the dataspace");
      check(o.layout_address, nbytes);
    }
    /* chunked (possibly deflated) data: no chunk can inflate by more than ~1000 */
    if (o.layout_class == 2 && nbytes / 4096 > size_) fail("dataset \"" + path + "\": a dataspace far larger than the file (corrupted?)");
    std::vector<double> out(n);
    if (o.layout_class == 1 || o.layout_class == 0) {
      if (o.layout_class == 1 && o.layout_address == UNDEFINED) return out; /* never written: the fill value (0) */
      convert(o.type, data_ + o.layout_address, n, out.data());
      return out;
    }
    /* chunked */
    const size_t nd = o.chunk_dims.size() - 1;
    if (nd != o.dims.size() || nd == 0 || nd > 2) fail("dataset \"" + path + "\": only 1-D and 2-D chunked datasets are read");
    uint64_t chunk_elements = 1;
    for (size_t k = 0; k < nd; ++k) chunk_elements = mul(chunk_elements, o.chunk_dims[k]);
    if (chunk_elements == 0 || mul(chunk_elements, esize) > (uint64_t(1) << 32))
      fail("dataset \"" + path + "\": chunks of zero or more than 4 GiB (corrupted?)");
    std::vector<uint8_t> raw(chunk_elements * esize), tmp;
    std::vector<double> cvals(chunk_elements);
    if (o.layout_address != UNDEFINED)
      read_chunks(o, o.layout_address, nd, chunk_elements, raw, tmp, cvals, out, 0, -1);
    return out;
  }

private:
  static constexpr uint64_t UNDEFINED = ~uint64_t(0);
  struct Datatype;
  struct Member {
    std::string name;
    uint32_t offset = 0;
    int cls = -1; /* the member's own (simple) type */
    uint32_t size = 0;
    bool is_signed = false;
  };
  struct Datatype {
    int cls = -1;
    uint32_t size = 0;
    bool is_signed = false;
    std::vector<Member> members; /* class 6 (compound, version 1) with simple members */
  };
  struct Attribute {
    std::string name;
    Datatype type;
    uint64_t count = 1, data = 0;
  };
  struct Object {
    uint64_t address = 0;
    bool is_group = false, has_layout = false, has_type = false, has_space = false;
    uint64_t btree = 0, heap = 0;
    Datatype type;
    std::vector<uint64_t> dims;
    int layout_class = -1;
    uint64_t layout_address = 0, layout_size = 0;
    std::vector<uint32_t> chunk_dims;
    std::vector<int> filters; /* in pipeline order */
    std::vector<Attribute> attributes;
  };

  [[noreturn]] void fail(const std::string &what) const {
    cmi_error("HDF5 file \"%s\": %s!", filename_.c_str(), what.c_str());
  }
  void check(uint64_t offset, uint64_t n) const {
    if (offset > size_ || n > size_ - offset) fail("a block points outside the file (truncated?)");
  }
  uint8_t u8(uint64_t o) const { check(o, 1); return data_[o]; }
  uint16_t u16(uint64_t o) const { check(o, 2); uint16_t v; memcpy(&v, data_ + o, 2); return v; }
  uint32_t u32(uint64_t o) const { check(o, 4); uint32_t v; memcpy(&v, data_ + o, 4); return v; }
  uint64_t u64(uint64_t o) const { check(o, 8); uint64_t v; memcpy(&v, data_ + o, 8); return v; }
  static uint64_t pad8(uint64_t n) { return (n + 7) & ~uint64_t(7); }
  /* products of sizes that come from the file: overflow is an error, not a wrap-around */
  uint64_t mul(uint64_t a, uint64_t b) const {
    uint64_t r;
    if (__builtin_mul_overflow(a, b, &r)) fail("sizes whose product overflows (corrupted?)");
    return r;
  }
  uint64_t element_count(const std::vector<uint64_t> &dims) const {
    uint64_t n = 1;
    for (uint64_t d : dims) n = mul(n, d);
    return n;
  }
  /* largest dataset this reader materialises (doubles): nothing the reference writes comes near, and a corrupted
   * dataspace must give an error, not std::bad_alloc */
  static constexpr uint64_t MAX_ELEMENTS = uint64_t(1) << 34;

  Datatype parse_datatype(uint64_t o) const {
    Datatype t;
    t.cls = u8(o) & 0x0f;
    t.size = u32(o + 4);
    const uint8_t bits0 = u8(o + 1);
    if (t.cls == 0 || t.cls == 1) {
      if (bits0 & 1) fail("big-endian data");
      t.is_signed = (bits0 & 8) != 0;
      if (t.cls == 1 && t.size != 8 && t.size != 4) fail("a floating point type that is not 4 or 8 bytes");
      if (t.cls == 0 && t.size != 1 && t.size != 2 && t.size != 4 && t.size != 8) fail("an unusual integer size");
    } else if (t.cls == 6) {
      /* compound, version 1: per member a name padded to 8 bytes, byte offset, dimensionality (+ 31 bytes of
       * array information), then the member's own datatype message */
      if ((u8(o) >> 4) != 1) fail("compound datatype version " + std::to_string(u8(o) >> 4));
      const uint16_t nmembers = u16(o + 1);
      uint64_t q = o + 8;
      for (uint16_t k = 0; k < nmembers; ++k) {
        Member m;
        check(q, 1);
        const char *name = reinterpret_cast<const char *>(data_ + q);
        const size_t len = strnlen(name, size_ - q);
        m.name.assign(name, len);
        q += pad8(len + 1);
        m.offset = u32(q);
        if (u8(q + 4) != 0) fail("array members of compound datatypes");
        q += 32;
        const Datatype mt = parse_datatype(q);
        if (mt.cls == 6) fail("nested compound datatypes");
        m.cls = mt.cls; m.size = mt.size; m.is_signed = mt.is_signed;
        if ((uint64_t)m.offset + m.size > t.size) fail("a compound member that does not fit in its record");
        q += 8 + (mt.cls == 0 ? 4 : (mt.cls == 1 ? 12 : 0));
        t.members.push_back(m);
      }
    } else if (t.cls != 3) {
      fail("datatype class " + std::to_string(t.cls) + " (only integers, floats, fixed-length strings and compounds of those are read)");
    }
    return t;
  }
  /* returns the dimensions; rank 0 = scalar */
  std::vector<uint64_t> parse_dataspace(uint64_t o) const {
    const uint8_t version = u8(o), rank = u8(o + 1);
    uint64_t p;
    if (version == 1) p = o + 8;
    else if (version == 2) p = o + 4;
    else fail("dataspace version " + std::to_string(version));
    std::vector<uint64_t> dims(rank);
    for (int k = 0; k < rank; ++k) dims[k] = u64(p + 8 * k);
    return dims;
  }
  void convert(const Datatype &t, const uint8_t *src, uint64_t n, double *out) const {
    if (t.cls == 1 && t.size == 8) { memcpy(out, src, 8 * n); return; }
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t *p = src + i * t.size;
      if (t.cls == 1) { float v; memcpy(&v, p, 4); out[i] = v; }
      else if (t.cls == 0) {
        if (t.size == 8) { if (t.is_signed) { int64_t v; memcpy(&v, p, 8); out[i] = (double)v; } else { uint64_t v; memcpy(&v, p, 8); out[i] = (double)v; } }
        else if (t.size == 4) { if (t.is_signed) { int32_t v; memcpy(&v, p, 4); out[i] = v; } else { uint32_t v; memcpy(&v, p, 4); out[i] = v; } }
        else if (t.size == 2) { if (t.is_signed) { int16_t v; memcpy(&v, p, 2); out[i] = v; } else { uint16_t v; memcpy(&v, p, 2); out[i] = v; } }
        else { out[i] = t.is_signed ? (double)(int8_t)p[0] : (double)p[0]; }
      } else {
        fail("a string where numbers are expected");
      }
    }
  }

  Object parse_object(uint64_t address) const {
    Object o;
    o.address = address;
    if (u8(address) != 1) fail("object header version " + std::to_string(u8(address)) + " (only version 1 is read)");
    const uint16_t nmsg = u16(address + 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{address + 16, u32(address + 8)}};
    uint16_t seen = 0;
    for (size_t ib = 0; ib < blocks.size() && seen < nmsg; ++ib) {
      uint64_t p = blocks[ib].first;
      const uint64_t end = p + blocks[ib].second;
      check(p, blocks[ib].second);
      while (p + 8 <= end && seen < nmsg) {
        const uint16_t type = u16(p), size = u16(p + 2);
        const uint8_t flags = u8(p + 4);
        const uint64_t body = p + 8;
        ++seen;
        /* bit 1: the body is a reference to a shared message elsewhere in the file, not the message itself */
        if ((flags & 2) && type != 0x0000) {
          if (type == 0x0001 || type == 0x0003 || type == 0x0008 || type == 0x000B || type == 0x0011)
            fail("shared object header messages");
          p = body + size; /* a shared attribute (or anything else) is skipped like an unknown one */
          continue;
        }
        switch (type) {
        case 0x0010: blocks.push_back({u64(body), u64(body + 8)}); break;
        case 0x0011: o.is_group = true; o.btree = u64(body); o.heap = u64(body + 8); break;
        case 0x0002: fail("new-style groups (link info messages)");
        case 0x0001: o.dims = parse_dataspace(body); o.has_space = true; break;
        case 0x0003: o.type = parse_datatype(body); o.has_type = true; break;
        case 0x0008: {
          if (u8(body) != 3) fail("data layout version " + std::to_string(u8(body)));
          o.layout_class = u8(body + 1);
          o.has_layout = true;
          if (o.layout_class == 1) { o.layout_address = u64(body + 2); o.layout_size = u64(body + 10); }
          else if (o.layout_class == 0) { o.layout_size = u16(body + 2); o.layout_address = body + 4; }
          else if (o.layout_class == 2) {
            const uint8_t nd = u8(body + 2);
            o.layout_address = u64(body + 3);
            o.chunk_dims.resize(nd);
            for (int k = 0; k < nd; ++k) o.chunk_dims[k] = u32(body + 11 + 4 * k);
          } else fail("data layout class " + std::to_string(o.layout_class));
          break;
        }
        case 0x000B: {
          if (u8(body) != 1) fail("filter pipeline version " + std::to_string(u8(body)));
          const uint8_t nfilters = u8(body + 1);
          uint64_t q = body + 8;
          for (int k = 0; k < nfilters; ++k) {
            const uint16_t id = u16(q), name_length = u16(q + 2), ncd = u16(q + 6);
            if (id != 1 && id != 2) fail("filter " + std::to_string(id) + " (only deflate and shuffle are read)");
            o.filters.push_back(id);
            q += 8 + pad8(name_length) + 4 * (uint64_t)(ncd + (ncd & 1));
          }
          break;
        }
        case 0x000C: {
          /* an attribute of a kind this reader does not know (compound, variable length, ...) is skipped */
          try {
            if (u8(body) != 1) fail("attribute message version " + std::to_string(u8(body)));
            const uint16_t name_size = u16(body + 2), type_size = u16(body + 4), space_size = u16(body + 6);
            Attribute a;
            uint64_t q = body + 8;
            check(q, name_size);
            a.name.assign(reinterpret_cast<const char *>(data_ + q), strnlen(reinterpret_cast<const char *>(data_ + q), name_size));
            q += pad8(name_size);
            a.type = parse_datatype(q);
            q += pad8(type_size);
            a.count = element_count(parse_dataspace(q));
            q += pad8(space_size);
            a.data = q;
            check(q, mul(a.count, a.type.size));
            o.attributes.push_back(a);
          } catch (const Error &) {
          }
          break;
        }
        default: break;
        }
        p = body + size;
      }
    }
    if (seen != nmsg) fail("an object header with missing messages");
    return o;
  }

  /* the links of a group: name -> object header address */
  std::map<std::string, uint64_t> links(const Object &g) const {
    std::map<std::string, uint64_t> out;
    if (!g.is_group) fail("a path component that is not a group");
    check(g.heap, 32);
    if (memcmp(data_ + g.heap, "HEAP", 4) != 0) fail("a bad local heap");
    const uint64_t segment = u64(g.heap + 24);
    if (g.btree != UNDEFINED) walk_group_node(g.btree, segment, out, 0);
    return out;
  }
  void walk_group_node(uint64_t node, uint64_t segment, std::map<std::string, uint64_t> &out, int depth) const {
    check(node, 8);
    if (depth > 16) fail("a group B-tree that is too deep");
    if (memcmp(data_ + node, "SNOD", 4) == 0) {
      const uint16_t nsym = u16(node + 6);
      for (uint16_t k = 0; k < nsym; ++k) {
        const uint64_t e = node + 8 + 40 * (uint64_t)k;
        const uint64_t name = segment + u64(e);
        check(name, 1);
        const char *s = reinterpret_cast<const char *>(data_ + name);
        out[std::string(s, strnlen(s, size_ - name))] = u64(e + 8);
      }
      return;
    }
    if (memcmp(data_ + node, "TREE", 4) != 0 || u8(node + 4) != 0) fail("a bad group B-tree node");
    const uint16_t used = u16(node + 6);
    for (uint16_t k = 0; k < used; ++k) walk_group_node(u64(node + 24 + 16 * (uint64_t)k + 8), segment, out, depth + 1);
  }
  bool find(const std::string &path, uint64_t &address) const {
    Object o = root_;
    address = root_.address;
    size_t pos = 0;
    while (pos < path.size()) {
      const size_t next = path.find('/', pos);
      const std::string part = path.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
      pos = (next == std::string::npos) ? path.size() : next + 1;
      if (part.empty()) continue;
      if (!o.is_group) return false;
      const auto l = links(o);
      const auto it = l.find(part);
      if (it == l.end()) return false;
      address = it->second;
      o = parse_object(address);
    }
    return true;
  }
  Object open_object(const std::string &path) const {
    uint64_t address;
    if (!find(path, address)) fail("\"" + path + "\" does not exist");
    return parse_object(address);
  }
  Attribute attribute(const std::string &path, const std::string &name) const {
    const Object o = open_object(path);
    for (const Attribute &a : o.attributes)
      if (a.name == name) return a;
    fail("attribute \"" + name + "\" of \"" + path + "\" does not exist");
  }

  void read_chunks(const Object &o, uint64_t node, size_t nd, uint64_t chunk_elements, std::vector<uint8_t> &raw,
                   std::vector<uint8_t> &tmp, std::vector<double> &cvals, std::vector<double> &out, int depth,
                   int expected_level) const {
    check(node, 24);
    if (memcmp(data_ + node, "TREE", 4) != 0 || u8(node + 4) != 1) fail("a bad chunk B-tree node");
    const uint8_t level = u8(node + 5);
    /* a child sits exactly one level below its parent: a node that points at itself or upwards is a cycle */
    if (depth > 16 || (expected_level >= 0 && level != expected_level))
      fail("a chunk B-tree that is too deep or cyclic");
    const uint16_t used = u16(node + 6);
    const uint64_t key_size = 8 + 8 * (nd + 1);
    const uint64_t esize = o.type.size;
    for (uint16_t k = 0; k < used; ++k) {
      const uint64_t key = node + 24 + (uint64_t)k * (key_size + 8);
      const uint64_t child = u64(key + key_size);
      if (level > 0) {
        read_chunks(o, child, nd, chunk_elements, raw, tmp, cvals, out, depth + 1, (int)level - 1);
        continue;
      }
      const uint32_t nbytes = u32(key), mask = u32(key + 4);
      uint64_t offset[2] = {u64(key + 8), nd > 1 ? u64(key + 16) : 0};
      check(child, nbytes);
      const uint8_t *src = data_ + child;
      uint64_t have = nbytes;
      /* undo the pipeline back to front; a set mask bit means that filter was skipped for this chunk */
      for (int f = (int)o.filters.size() - 1; f >= 0; --f) {
        if (mask & (1u << f)) continue;
        if (o.filters[f] == 1) {
          uLongf n = raw.size();
          tmp.assign(src, src + have);
          if (uncompress(raw.data(), &n, tmp.data(), have) != Z_OK) fail("a chunk that does not inflate");
          src = raw.data();
          have = n;
        } else { /* shuffle: byte b of element i sits at b * count + i */
          const uint64_t count = have / esize;
          tmp.assign(src, src + have);
          if (raw.size() < have) fail("a chunk larger than its dimensions");
          for (uint64_t i = 0; i < count; ++i)
            for (uint64_t b = 0; b < esize; ++b) raw[i * esize + b] = tmp[b * count + i];
          src = raw.data();
        }
      }
      if (have != chunk_elements * esize) fail("a chunk whose size does not match its dimensions");
      convert(o.type, src, chunk_elements, cvals.data());
      /* chunk offsets come from the file: they must lie inside the dataspace (no wrap-around in the index) */
      if (offset[0] >= o.dims[0] || (nd > 1 && offset[1] >= o.dims[1])) fail("a chunk outside its dataspace");
      if (nd == 1) {
        for (uint64_t i = 0; i < o.chunk_dims[0] && offset[0] + i < o.dims[0]; ++i) out.at(offset[0] + i) = cvals[i];
      } else {
        for (uint64_t i = 0; i < o.chunk_dims[0] && offset[0] + i < o.dims[0]; ++i)
          for (uint64_t j = 0; j < o.chunk_dims[1] && offset[1] + j < o.dims[1]; ++j)
            out.at((offset[0] + i) * o.dims[1] + offset[1] + j) = cvals[i * o.chunk_dims[1] + j];
      }
    }
  }

  std::string filename_;
  const uint8_t *data_ = nullptr;
  uint64_t size_ = 0;
  Object root_;
};

} // namespace hdf5
} // namespace cmi
