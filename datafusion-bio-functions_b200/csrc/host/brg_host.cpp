// brg_host.cpp -- host-side operators above the C ABI (see include/bioranges_host.h).
//
// The reference's operators are Rust DataFusion plan nodes / table providers.  There is no Rust toolchain in this image,
// so the part of them that sits between "Arrow record batches" and "the interval-overlap hot path" is restated here in
// C++: argument grammar, column resolution and its errors, contig dictionary, build-once / probe-per-batch streaming,
// output schemas, `take` of the payload columns.  All interval work goes through the brg_* C ABI (CUDA); nothing here
// computes overlaps on the CPU.
//
// Citations are relative to /root/reference/datafusion/bio-function-ranges/src.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../../include/bioranges_host.h"

// ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---------------------------------
extern "C" {
struct ArrowSchema {
    const char* format;
    const char* name;
    const char* metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema** children;
    struct ArrowSchema* dictionary;
    void (*release)(struct ArrowSchema*);
    void* private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void** buffers;
    struct ArrowArray** children;
    struct ArrowArray* dictionary;
    void (*release)(struct ArrowArray*);
    void* private_data;
};
}
static const int64_t kArrowNullable = 2;

namespace brg {
void set_error(const std::string& msg);  // brg_api.cu (thread-local text behind brg_last_error)
}

namespace {

// ---- errors --------------------------------------------------------------------------------------------------------------
struct Err {
    int code;
    std::string msg;
};
#define BRGH_TRY(expr)                   \
    do {                                 \
        Err _e = (expr);                 \
        if (_e.code) return _e;          \
    } while (0)
static const Err kOk{0, ""};
static Err plan(std::string m) { return Err{BRGH_ERR_PLAN, std::move(m)}; }
static Err exec(std::string m) { return Err{BRGH_ERR_EXECUTION, std::move(m)}; }
static Err notimpl(std::string m) { return Err{BRGH_ERR_NOT_IMPLEMENTED, std::move(m)}; }
static Err internal(std::string m) { return Err{BRGH_ERR_INTERNAL, std::move(m)}; }
static int finish(const Err& e) {
    if (e.code) brg::set_error(e.msg);
    return e.code;
}
// failure of a brg_* call: its message is already in place
static Err from_abi(int rc) { return rc ? Err{rc, brg_last_error()} : kOk; }

static std::string lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// ---- columns, batches, tables --------------------------------------------------------------------------------------------
struct ArrayHolder {  // owns an imported struct array (and through it every child buffer)
    ArrowArray arr{};
    ~ArrayHolder() {
        if (arr.release) arr.release(&arr);
    }
};
struct OwnedBuffers {  // buffers allocated here
    std::vector<void*> ptrs;
    ~OwnedBuffers() {
        for (void* p : ptrs) free(p);
    }
    void* alloc(size_t bytes) {
        void* p = nullptr;
        if (posix_memalign(&p, 64, std::max<size_t>(bytes, 64))) return nullptr;
        ptrs.push_back(p);
        return p;
    }
};

struct Field {
    std::string name, format;
    bool nullable = true;
};
struct Column {
    Field f;
    int64_t length = 0, offset = 0, null_count = 0;
    std::vector<const void*> buffers;
    std::shared_ptr<void> keep;
    const uint8_t* validity() const { return buffers.empty() ? nullptr : (const uint8_t*)buffers[0]; }
    bool is_null(int64_t i) const {
        const uint8_t* v = validity();
        if (!v || null_count == 0) return false;
        const int64_t b = offset + i;
        return !((v[b >> 3] >> (b & 7)) & 1);
    }
};
struct Batch {
    std::vector<Column> cols;
    int64_t rows = 0;
    const Column* find(const std::string& name) const {
        for (auto& c : cols)
            if (c.f.name == name) return &c;
        return nullptr;
    }
    std::string names() const {  // {:?} of Vec<&String>
        std::string s = "[";
        for (size_t i = 0; i < cols.size(); i++) s += (i ? ", \"" : "\"") + cols[i].f.name + "\"";
        return s + "]";
    }
};
struct Table {
    std::vector<Field> schema;
    std::vector<Batch> batches;
    int64_t rows() const {
        int64_t n = 0;
        for (auto& b : batches) n += b.rows;
        return n;
    }
};

// Rust's `{:?}` of the arrow DataType for the formats the tests can produce
static std::string type_name(const std::string& f) {
    static const std::map<std::string, std::string> m = {
        {"n", "Null"}, {"b", "Boolean"}, {"c", "Int8"}, {"C", "UInt8"}, {"s", "Int16"}, {"S", "UInt16"}, {"i", "Int32"},
        {"I", "UInt32"}, {"l", "Int64"}, {"L", "UInt64"}, {"e", "Float16"}, {"f", "Float32"}, {"g", "Float64"},
        {"u", "Utf8"}, {"U", "LargeUtf8"}, {"vu", "Utf8View"}, {"z", "Binary"}, {"Z", "LargeBinary"}, {"tdD", "Date32"},
        {"tdm", "Date64"}};
    auto it = m.find(f);
    return it == m.end() ? "DataType(" + f + ")" : it->second;
}
static int fixed_width(const std::string& f) {
    if (f == "c" || f == "C") return 1;
    if (f == "s" || f == "S" || f == "e") return 2;
    if (f == "i" || f == "I" || f == "f" || f == "tdD") return 4;
    if (f == "l" || f == "L" || f == "g" || f == "tdm") return 8;
    if (f.rfind("ts", 0) == 0 || f.rfind("tD", 0) == 0) return 8;             // timestamp / duration
    if (f == "tts" || f == "ttm") return 4;
    if (f == "ttu" || f == "ttn") return 8;
    return 0;
}

static Err import_batch(ArrowArray* arr, ArrowSchema* sch, Batch* out, std::vector<Field>* schema_out) {
    if (!arr || !sch || !arr->release || !sch->release) return internal("record batch or schema already released");
    if (!sch->format || strcmp(sch->format, "+s") != 0) return internal("a record batch must be exported as a struct array");
    if (arr->n_children != sch->n_children) return internal("record batch and schema disagree on the number of columns");
    auto holder = std::make_shared<ArrayHolder>();
    holder->arr = *arr;       // move
    arr->release = nullptr;
    Batch b;
    b.rows = holder->arr.length;
    for (int64_t i = 0; i < sch->n_children; i++) {
        const ArrowSchema* cs = sch->children[i];
        const ArrowArray* ca = holder->arr.children[i];
        Column c;
        c.f.name = cs->name ? cs->name : "";
        c.f.format = cs->format ? cs->format : "";
        c.f.nullable = (cs->flags & kArrowNullable) != 0;
        // a sliced struct array: children carry their own offset, the parent's offset adds to it
        c.length = holder->arr.length;
        c.offset = ca->offset + holder->arr.offset;
        c.null_count = ca->null_count < 0 ? -1 : ca->null_count;
        for (int64_t k = 0; k < ca->n_buffers; k++) c.buffers.push_back(ca->buffers[k]);
        if (c.null_count < 0) {  // unknown: count
            int64_t n = 0;
            c.null_count = 1;
            for (int64_t r = 0; r < c.length; r++) n += c.is_null(r) ? 1 : 0;
            c.null_count = n;
        }
        c.keep = holder;
        b.cols.push_back(std::move(c));
    }
    if (schema_out) {
        schema_out->clear();
        for (auto& c : b.cols) schema_out->push_back(c.f);
    }
    sch->release(sch);  // the schema was moved to us as well
    *out = std::move(b);
    return kOk;
}

// ---- export --------------------------------------------------------------------------------------------------------------
struct ChildPriv {
    std::shared_ptr<void> keep;
    std::vector<const void*> buffers;
};
static void release_child(ArrowArray* a) {
    delete (ChildPriv*)a->private_data;
    a->release = nullptr;
}
struct ParentPriv {
    std::vector<ArrowArray> children;
    std::vector<ArrowArray*> child_ptrs;
    const void* no_buffer = nullptr;
};
static void release_parent(ArrowArray* a) {
    auto* p = (ParentPriv*)a->private_data;
    for (auto& c : p->children)
        if (c.release) c.release(&c);
    delete p;
    a->release = nullptr;
}
static void export_batch(const Batch& b, ArrowArray* out) {
    auto* pp = new ParentPriv();
    pp->children.resize(b.cols.size());
    for (size_t i = 0; i < b.cols.size(); i++) {
        const Column& c = b.cols[i];
        auto* cp = new ChildPriv();
        cp->keep = c.keep;
        cp->buffers = c.buffers;
        ArrowArray& a = pp->children[i];
        memset(&a, 0, sizeof(a));
        a.length = c.length;
        a.null_count = c.null_count;
        a.offset = c.offset;
        a.n_buffers = (int64_t)cp->buffers.size();
        a.buffers = cp->buffers.data();
        a.release = release_child;
        a.private_data = cp;
    }
    for (auto& c : pp->children) pp->child_ptrs.push_back(&c);
    memset(out, 0, sizeof(*out));
    out->length = b.rows;
    out->n_buffers = 1;
    out->buffers = &pp->no_buffer;
    out->n_children = (int64_t)pp->children.size();
    out->children = pp->child_ptrs.data();
    out->release = release_parent;
    out->private_data = pp;
}
struct SchemaPriv {
    std::string format, name;
    std::vector<ArrowSchema> children;
    std::vector<ArrowSchema*> child_ptrs;
};
static void release_schema(ArrowSchema* s) {
    auto* p = (SchemaPriv*)s->private_data;
    for (auto& c : p->children)
        if (c.release) c.release(&c);
    delete p;
    s->release = nullptr;
}
static void fill_schema(ArrowSchema* s, const std::string& format, const std::string& name, bool nullable,
                        const std::vector<Field>* children) {
    auto* p = new SchemaPriv();
    p->format = format;
    p->name = name;
    memset(s, 0, sizeof(*s));
    if (children) {
        p->children.resize(children->size());
        for (size_t i = 0; i < children->size(); i++)
            fill_schema(&p->children[i], (*children)[i].format, (*children)[i].name, (*children)[i].nullable, nullptr);
        for (auto& c : p->children) p->child_ptrs.push_back(&c);
        s->n_children = (int64_t)p->children.size();
        s->children = p->child_ptrs.data();
    }
    s->format = p->format.c_str();
    s->name = p->name.c_str();
    s->flags = nullable ? kArrowNullable : 0;
    s->release = release_schema;
    s->private_data = p;
}
static void export_schema(const Batch& b, ArrowSchema* out) {
    std::vector<Field> fs;
    for (auto& c : b.cols) fs.push_back(c.f);
    fill_schema(out, "+s", "", false, &fs);
}

// ---- new columns ---------------------------------------------------------------------------------------------------------
template <typename T>
static Column make_primitive(const std::string& name, const char* format, bool nullable, std::vector<T>&& values,
                             const uint8_t* valid_bits /*may be NULL*/, int64_t null_count) {
    auto ob = std::make_shared<OwnedBuffers>();
    const size_t n = values.size();
    T* data = (T*)ob->alloc(n * sizeof(T));
    if (n) memcpy(data, values.data(), n * sizeof(T));
    uint8_t* bits = nullptr;
    if (valid_bits && null_count > 0) {
        bits = (uint8_t*)ob->alloc((n + 7) / 8);
        memcpy(bits, valid_bits, (n + 7) / 8);
    }
    Column c;
    c.f = Field{name, format, nullable};
    c.length = (int64_t)n;
    c.null_count = bits ? null_count : 0;
    c.buffers = {bits, data};
    c.keep = ob;
    return c;
}

// ---- column access (array_utils.rs) ----------------------------------------------------------------------------------------
struct ContigArray {  // Utf8 / LargeUtf8 / Utf8View (array_utils.rs:10-24)
    const Column* c = nullptr;
    int kind = 0;  // 0 Utf8, 1 LargeUtf8, 2 Utf8View
    std::string_view value(int64_t i) const {
        const int64_t r = c->offset + i;
        if (kind == 0) {
            const int32_t* o = (const int32_t*)c->buffers[1];
            return std::string_view((const char*)c->buffers[2] + o[r], (size_t)(o[r + 1] - o[r]));
        }
        if (kind == 1) {
            const int64_t* o = (const int64_t*)c->buffers[1];
            return std::string_view((const char*)c->buffers[2] + o[r], (size_t)(o[r + 1] - o[r]));
        }
        const uint8_t* v = (const uint8_t*)c->buffers[1] + 16 * r;  // 16-byte view
        int32_t len, buf, off;
        memcpy(&len, v, 4);
        if (len <= 12) return std::string_view((const char*)v + 4, (size_t)len);
        memcpy(&buf, v + 8, 4);
        memcpy(&off, v + 12, 4);
        return std::string_view((const char*)c->buffers[2 + buf] + off, (size_t)len);
    }
};

static Err contig_array(const Batch& b, const std::string& name, ContigArray* out) {
    const Column* c = b.find(name);
    if (!c) return plan("contig column '" + name + "' not found in batch with columns: " + b.names());  // array_utils.rs:186-197
    if (c->f.format == "u") *out = ContigArray{c, 0};
    else if (c->f.format == "U") *out = ContigArray{c, 1};
    else if (c->f.format == "vu") *out = ContigArray{c, 2};
    else
        return notimpl("unsupported data type " + type_name(c->f.format) + " for contig column '" + name +
                       "'; expected Utf8, LargeUtf8, or Utf8View");  // array_utils.rs:236-241
    return kOk;
}

// extract_pos_array + PosArray::resolve (array_utils.rs:73-131, 246-300)
static Err resolve_positions(const Batch& b, const std::string& name, const char* label, std::vector<int32_t>* out,
                             bool cast_error_style = false) {
    const Column* c = b.find(name);
    if (!c) return plan(std::string(label) + " column '" + name + "' not found in batch with columns: " + b.names());
    const std::string& f = c->f.format;
    if (f != "i" && f != "l" && f != "I" && f != "L")
        return notimpl("unsupported data type " + type_name(f) + " for " + label + " column '" + name +
                       "'; expected Int32, Int64, UInt32, or UInt64");
    if (c->null_count > 0)
        return exec("coordinate column contains null values; nearest requires non-null coordinates");  // array_utils.rs:80-85
    const size_t base = out->size();
    out->resize(base + (size_t)c->length);
    int32_t* dst = out->data() + base;
    const int64_t n = c->length;
    auto overflow = [&](const std::string& v, int64_t i) {
        if (cast_error_style)  // the join casts with arrow (interval_join.rs:1697-1709; text pinned by interval_join.rs:1994-1997)
            return Err{BRGH_ERR_ARROW, "Arrow error: Cast error: Can't cast value " + v + " to type Int32"};
        return exec("coordinate value " + v + " at row " + std::to_string(i) + " overflows i32 (max 2147483647)");
    };
    if (f == "i") {
        if (n) memcpy(dst, (const int32_t*)c->buffers[1] + c->offset, (size_t)n * 4);
    } else if (f == "l") {
        const int64_t* s = (const int64_t*)c->buffers[1] + c->offset;
        for (int64_t i = 0; i < n; i++) {
            if (s[i] < INT32_MIN || s[i] > INT32_MAX) return overflow(std::to_string(s[i]), i);
            dst[i] = (int32_t)s[i];
        }
    } else if (f == "I") {
        const uint32_t* s = (const uint32_t*)c->buffers[1] + c->offset;
        for (int64_t i = 0; i < n; i++) {
            if (s[i] > (uint32_t)INT32_MAX) return overflow(std::to_string(s[i]), i);
            dst[i] = (int32_t)s[i];
        }
    } else {
        const uint64_t* s = (const uint64_t*)c->buffers[1] + c->offset;
        for (int64_t i = 0; i < n; i++) {
            if (s[i] > (uint64_t)INT32_MAX) return overflow(std::to_string(s[i]), i);
            dst[i] = (int32_t)s[i];
        }
    }
    return kOk;
}

// Arrow ingest: the value buffer of a position column as it is, for the *_arrow entry points that narrow
// and check on the device.  Same type / NULL errors as PosArray::resolve (array_utils.rs:73-98); the bytes of the column's
// rows are appended to `raw` so that coalesced batches form one buffer.
static Err position_bytes(const Batch& b, const std::string& name, const char* label, std::vector<uint8_t>* raw, int* pos_type) {
    const Column* c = b.find(name);
    if (!c) return plan(std::string(label) + " column '" + name + "' not found in batch with columns: " + b.names());
    const std::string& f = c->f.format;
    if (f != "i" && f != "l" && f != "I" && f != "L")
        return notimpl("unsupported data type " + type_name(f) + " for " + label + " column '" + name +
                       "'; expected Int32, Int64, UInt32, or UInt64");
    if (c->null_count > 0)
        return exec("coordinate column contains null values; nearest requires non-null coordinates");  // array_utils.rs:80-85
    *pos_type = f == "i" ? BRG_POS_INT32 : f == "l" ? BRG_POS_INT64 : f == "I" ? BRG_POS_UINT32 : BRG_POS_UINT64;
    const size_t w = (f == "l" || f == "L") ? 8 : 4;
    const uint8_t* src = (const uint8_t*)c->buffers[1] + (size_t)c->offset * w;
    if (c->length) raw->insert(raw->end(), src, src + (size_t)c->length * w);
    return kOk;
}
// the reference resolves batch by batch, start column first: of the first offending rows the device reports for a coalesced
// group, the one in the earlier batch wins (start on a tie) and is reported with its batch-local row index
static Err ingest_overflow(const uint64_t bad[2], const std::vector<uint64_t>& batch_first_row, const std::vector<uint8_t>& sraw,
                           int s_type, const std::vector<uint8_t>& eraw, int e_type) {
    auto batch_of = [&](uint64_t row) {
        size_t bi = std::upper_bound(batch_first_row.begin(), batch_first_row.end(), row) - batch_first_row.begin();
        return bi - 1;
    };
    const bool has_s = bad[0] != ~0ull, has_e = bad[1] != ~0ull;
    const bool use_s = has_s && (!has_e || batch_of(bad[0]) <= batch_of(bad[1]));
    const uint64_t row = use_s ? bad[0] : bad[1];
    const int t = use_s ? s_type : e_type;
    const uint8_t* col = use_s ? sraw.data() : eraw.data();
    std::string v;
    if (t == BRG_POS_INT64) v = std::to_string(reinterpret_cast<const int64_t*>(col)[row]);
    else if (t == BRG_POS_UINT32) v = std::to_string(reinterpret_cast<const uint32_t*>(col)[row]);
    else v = std::to_string(reinterpret_cast<const uint64_t*>(col)[row]);
    return exec("coordinate value " + v + " at row " + std::to_string(row - batch_first_row[batch_of(row)]) + " overflows i32 (max 2147483647)");
}

// PosArray::resolve_i64 (array_utils.rs:137-171): Int32 / Int64 / UInt32 widen, UInt64 is checked against i64::MAX
static Err resolve_positions_i64(const Batch& b, const std::string& name, const char* label, std::vector<int64_t>* out) {
    const Column* c = b.find(name);
    if (!c) return plan(std::string(label) + " column '" + name + "' not found in batch with columns: " + b.names());
    const std::string& f = c->f.format;
    if (f != "i" && f != "l" && f != "I" && f != "L")
        return notimpl("unsupported data type " + type_name(f) + " for " + label + " column '" + name +
                       "'; expected Int32, Int64, UInt32, or UInt64");
    if (c->null_count > 0) return exec("coordinate column contains null values; requires non-null coordinates");
    const int64_t n = c->length;
    const size_t base = out->size();
    out->resize(base + (size_t)n);
    int64_t* dst = out->data() + base;
    if (f == "i") { const int32_t* s = (const int32_t*)c->buffers[1] + c->offset; for (int64_t i = 0; i < n; i++) dst[i] = s[i]; }
    else if (f == "l") { if (n) memcpy(dst, (const int64_t*)c->buffers[1] + c->offset, (size_t)n * 8); }
    else if (f == "I") { const uint32_t* s = (const uint32_t*)c->buffers[1] + c->offset; for (int64_t i = 0; i < n; i++) dst[i] = s[i]; }
    else {
        const uint64_t* s = (const uint64_t*)c->buffers[1] + c->offset;
        for (int64_t i = 0; i < n; i++) {
            if (s[i] > (uint64_t)INT64_MAX)
                return exec("coordinate value " + std::to_string(s[i]) + " at row " + std::to_string(i) + " overflows i64 (max 9223372036854775807)");
            dst[i] = (int64_t)s[i];
        }
    }
    return kOk;
}

// ---- take (arrow::compute::take over a chunked source) ------------------------------------------------------------------------
// idx[i] is a row of the concatenation of `src` batches; idx_valid (LSB-first bitmap, may be NULL) marks NULL indices
static Err take_column(const std::vector<const Column*>& src, const std::vector<int64_t>& batch_first_row, const uint32_t* idx,
                       const uint8_t* idx_valid, int64_t n, const Field& out_field, Column* out) {
    const std::string& f = out_field.format;
    auto ob = std::make_shared<OwnedBuffers>();
    uint8_t* bits = (uint8_t*)ob->alloc((size_t)(n + 7) / 8);
    memset(bits, 0, (size_t)(n + 7) / 8);
    int64_t nulls = 0;
    const size_t nb = src.size();
    auto locate = [&](uint32_t g, size_t* b, int64_t* r) {
        size_t lo = 0, hi = nb;  // last batch whose first row <= g
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (batch_first_row[mid] <= (int64_t)g) lo = mid; else hi = mid;
        }
        *b = lo;
        *r = (int64_t)g - batch_first_row[lo];
    };
    auto index_null = [&](int64_t i) { return idx_valid && !((idx_valid[i >> 3] >> (i & 7)) & 1); };
    Column c;
    c.f = out_field;
    c.length = n;
    const int w = fixed_width(f);
    if (w) {
        uint8_t* data = (uint8_t*)ob->alloc((size_t)n * w);
        for (int64_t i = 0; i < n; i++) {
            size_t b; int64_t r;
            bool null = index_null(i) || nb == 0;
            if (!null) { locate(idx[i], &b, &r); null = src[b]->is_null(r); }
            if (null) { memset(data + i * w, 0, w); nulls++; continue; }
            memcpy(data + i * w, (const uint8_t*)src[b]->buffers[1] + (src[b]->offset + r) * w, w);
            bits[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        c.buffers = {bits, data};
    } else if (f == "b") {
        uint8_t* data = (uint8_t*)ob->alloc((size_t)(n + 7) / 8);
        memset(data, 0, (size_t)(n + 7) / 8);
        for (int64_t i = 0; i < n; i++) {
            size_t b; int64_t r;
            bool null = index_null(i) || nb == 0;
            if (!null) { locate(idx[i], &b, &r); null = src[b]->is_null(r); }
            if (null) { nulls++; continue; }
            const int64_t p = src[b]->offset + r;
            if ((((const uint8_t*)src[b]->buffers[1])[p >> 3] >> (p & 7)) & 1) data[i >> 3] |= (uint8_t)(1u << (i & 7));
            bits[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        c.buffers = {bits, data};
    } else if (f == "u" || f == "U" || f == "z" || f == "Z") {
        const bool large = (f == "U" || f == "Z");
        std::vector<std::pair<const char*, int64_t>> parts((size_t)n, {nullptr, 0});
        int64_t total = 0;
        for (int64_t i = 0; i < n; i++) {
            size_t b; int64_t r;
            bool null = index_null(i) || nb == 0;
            if (!null) { locate(idx[i], &b, &r); null = src[b]->is_null(r); }
            if (null) { nulls++; continue; }
            const int64_t p = src[b]->offset + r;
            int64_t o0, o1;
            if (large) { const int64_t* o = (const int64_t*)src[b]->buffers[1]; o0 = o[p]; o1 = o[p + 1]; }
            else { const int32_t* o = (const int32_t*)src[b]->buffers[1]; o0 = o[p]; o1 = o[p + 1]; }
            parts[(size_t)i] = {(const char*)src[b]->buffers[2] + o0, o1 - o0};
            total += o1 - o0;
            bits[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        if (!large && total > INT32_MAX) return Err{BRGH_ERR_ARROW, "Arrow error: offset overflow taking a Utf8 column"};
        char* data = (char*)ob->alloc((size_t)total);
        void* offs = ob->alloc((size_t)(n + 1) * (large ? 8 : 4));
        int64_t pos = 0;
        for (int64_t i = 0; i <= n; i++) {
            if (large) ((int64_t*)offs)[i] = pos; else ((int32_t*)offs)[i] = (int32_t)pos;
            if (i < n && parts[(size_t)i].second) { memcpy(data + pos, parts[(size_t)i].first, (size_t)parts[(size_t)i].second); pos += parts[(size_t)i].second; }
        }
        c.buffers = {bits, offs, data};
    } else {
        return notimpl("take of a " + type_name(f) + " column ('" + out_field.name + "') is not supported by the host layer");
    }
    c.null_count = nulls;
    if (nulls == 0) c.buffers[0] = nullptr;
    c.keep = ob;
    *out = std::move(c);
    return kOk;
}

// take of every column of a table; output fields are `prefix + name`, nullable as given
static Err take_table(const Table& t, const uint32_t* idx, const uint8_t* idx_valid, int64_t n, const std::string& prefix,
                      int force_nullable /* -1 keep */, std::vector<Column>* out) {
    std::vector<int64_t> first;
    int64_t acc = 0;
    for (auto& b : t.batches) { first.push_back(acc); acc += b.rows; }
    for (size_t ci = 0; ci < t.schema.size(); ci++) {
        std::vector<const Column*> src;
        for (auto& b : t.batches) src.push_back(&b.cols[ci]);
        Field f = t.schema[ci];
        f.name = prefix + f.name;
        if (force_nullable >= 0) f.nullable = force_nullable != 0;
        Column c;
        BRGH_TRY(take_column(src, first, idx, idx_valid, n, f, &c));
        if (c.null_count > 0) c.f.nullable = true;
        out->push_back(std::move(c));
    }
    return kOk;
}

// ---- session ----------------------------------------------------------------------------------------------------------------
enum Algorithm { Coitrees, IntervalTree, ArrayIntervalTree, Lapper, SuperIntervals, CoitreesNearest, CoitreesCountOverlaps };
static const char* kAlgNames[] = {"Coitrees", "IntervalTree", "ArrayIntervalTree", "Lapper", "SuperIntervals",
                                  "CoitreesNearest", "CoitreesCountOverlaps"};

}  // namespace

struct brgh_session {
    int device = 0;
    brg_ctx* ctx = nullptr;
    // BioConfig (session_context.rs:48-58)
    bool prefer_interval_join = true;
    Algorithm algorithm = Coitrees;
    bool low_memory = false;
    uint64_t coalesce_rows = 8ull << 20;
    std::map<std::string, Table> tables;
};
struct brgh_result {
    std::vector<Batch> batches;
};

namespace {

static Err session_ctx(brgh_session* s, brg_ctx** out) {
    if (!s->ctx) BRGH_TRY(from_abi(brg_ctx_create(s->device, &s->ctx)));
    *out = s->ctx;
    return kOk;
}
static Err find_table(brgh_session* s, const char* name, const Table** out) {
    auto it = s->tables.find(name ? name : "");
    if (it == s->tables.end()) return plan(std::string("table '") + (name ? name : "") + "' not found");
    *out = &it->second;
    return kOk;
}
static Err parse_bool(const std::string& v, bool* out) {
    const std::string l = lower(v);
    if (l == "true" || l == "1") *out = true;
    else if (l == "false" || l == "0") *out = false;
    else return Err{BRGH_ERR_EXTERNAL, "Error parsing '" + v + "' as bool"};
    return kOk;
}

// ---- argument grammar of the table functions (table_function.rs:59-189) ----------------------------------------------------
struct ColTriple { std::string contig, start, end; };
struct UdtfArgs {
    ColTriple left, right;
    bool strict = false;
    int output_mode = BRGH_OVERLAP_JOIN;
};
static const char* kDefaultCols[3] = {"contig", "pos_start", "pos_end"};

static Err check_name(const std::string& v, const char* what, const std::string& fn) {
    if (v.find('`') != std::string::npos)
        return plan(fn + "() " + what + " must not contain backtick characters, got: " + v);  // table_function.rs:46-51
    return kOk;
}
static Err parse_col_triples(const std::vector<std::string>& a, const std::string& fn, UdtfArgs* out) {
    auto dflt = ColTriple{kDefaultCols[0], kDefaultCols[1], kDefaultCols[2]};
    if (a.empty()) { out->left = out->right = dflt; return kOk; }
    if (a.size() == 3) {
        for (auto& v : a) BRGH_TRY(check_name(v, "column name", fn));
        out->left = out->right = ColTriple{a[0], a[1], a[2]};
        return kOk;
    }
    if (a.size() == 6) {
        for (int i = 0; i < 3; i++) BRGH_TRY(check_name(a[i], "left column name", fn));
        for (int i = 3; i < 6; i++) BRGH_TRY(check_name(a[i], "right column name", fn));
        out->left = ColTriple{a[0], a[1], a[2]};
        out->right = ColTriple{a[3], a[4], a[5]};
        return kOk;
    }
    return plan(fn + "() expects 0, 3, or 6 column name arguments (got " + std::to_string(a.size()) + "). Usage: " + fn +
                "('left_table', 'right_table' [, col1, col2, col3 [, col4, col5, col6]] [, 'strict'|'weak'])");
}
// parse_col_args_extra (table_function.rs:72-96): a trailing 'strict' / 'weak' is the filter op
static Err parse_col_args(const char* const* args, int n, const std::string& fn, UdtfArgs* out) {
    std::vector<std::string> a;
    for (int i = 0; i < n; i++) a.push_back(args[i] ? args[i] : "");
    if (!a.empty()) {
        const std::string l = lower(a.back());
        if (l == "strict") { out->strict = true; a.pop_back(); }
        else if (l == "weak") { a.pop_back(); }
    }
    return parse_col_triples(a, fn, out);
}
// parse_overlap_args (table_function.rs:98-162): trailing filter op and output mode in any order, each at most once
static Err parse_overlap_args(const char* const* args, int n, const std::string& fn, UdtfArgs* out) {
    std::vector<std::string> a;
    for (int i = 0; i < n; i++) a.push_back(args[i] ? args[i] : "");
    bool has_filter = false, has_mode = false;
    while (!(a.size() == 0 || a.size() == 3 || a.size() == 6)) {
        const std::string l = lower(a.back());
        if ((l == "strict" || l == "weak") && !has_filter) { out->strict = (l == "strict"); has_filter = true; a.pop_back(); }
        else if ((l == "left" || l == "left_distinct") && !has_mode) { out->output_mode = BRGH_OVERLAP_LEFT_DISTINCT; has_mode = true; a.pop_back(); }
        else if ((l == "left_all" || l == "left_multiple") && !has_mode) { out->output_mode = BRGH_OVERLAP_LEFT; has_mode = true; a.pop_back(); }
        else if (l == "join" && !has_mode) { out->output_mode = BRGH_OVERLAP_JOIN; has_mode = true; a.pop_back(); }
        else break;
    }
    return parse_col_triples(a, fn, out);
}

// ---- build side ---------------------------------------------------------------------------------------------------------------
struct KeyDict {  // contig / join-key tuple -> dense id, first-seen order (interval_tree.rs:90-98; interval_join.rs:903-931)
    std::unordered_map<std::string, uint32_t> ids;
    uint32_t get_or_add(std::string_view k) {
        auto it = ids.find(std::string(k));
        if (it != ids.end()) return it->second;
        const uint32_t id = (uint32_t)ids.size();
        ids.emplace(std::string(k), id);
        return id;
    }
    uint32_t lookup(std::string_view k) const {
        auto it = ids.find(std::string(k));
        return it == ids.end() ? BRG_KEY_ABSENT : it->second;
    }
};

struct IndexGuard {
    brg_index* ix = nullptr;
    ~IndexGuard() { if (ix) brg_index_destroy(ix); }
};

// resolves (contig, start, end) of every batch of the build table and builds the device index
static Err build_from_table(brgh_session* s, const Table& t, const ColTriple& cols, uint32_t want, KeyDict* dict,
                            IndexGuard* out) {
    std::vector<uint32_t> key;
    std::vector<int32_t> st, en;
    for (auto& b : t.batches) {
        ContigArray ca;
        BRGH_TRY(contig_array(b, cols.contig, &ca));
        BRGH_TRY(resolve_positions(b, cols.start, "start", &st));
        BRGH_TRY(resolve_positions(b, cols.end, "end", &en));
        std::string_view cached;
        uint32_t cached_id = BRG_KEY_ABSENT;
        bool have = false;
        for (int64_t i = 0; i < b.rows; i++) {
            const std::string_view v = ca.value(i);
            if (!have || v != cached) { cached = v; cached_id = dict->get_or_add(v); have = true; }
            key.push_back(cached_id);
        }
    }
    brg_ctx* ctx;
    BRGH_TRY(session_ctx(s, &ctx));
    const uint32_t n_keys = std::max<uint32_t>(1, (uint32_t)dict->ids.size());
    return from_abi(brg_build(ctx, key.data(), st.data(), en.data(), key.size(), n_keys, want, &out->ix));
}

// run-length key column of one probe batch; `nearest_quirk`: the reference's nearest loop starts its contig cache at ""
// with no index, so a leading run of empty-string contigs never consults the map (nearest.rs:323-324,346-352)
static void probe_key_runs(const ContigArray& ca, int64_t rows, const KeyDict& dict, bool nearest_quirk, uint64_t row_base,
                           std::vector<uint32_t>* run_key, std::vector<uint64_t>* run_end) {
    std::string_view cached = "";
    uint32_t cached_id = BRG_KEY_ABSENT;
    bool have = nearest_quirk;  // Option<&str> = None for count/coverage (interval_tree.rs:246-247)
    for (int64_t i = 0; i < rows; i++) {
        const std::string_view v = ca.value(i);
        if (!have || v != cached) { cached = v; cached_id = dict.lookup(v); have = true; }
        if (!run_key->empty() && run_key->back() == cached_id && run_end->back() == row_base + (uint64_t)i) run_end->back()++;
        else { run_key->push_back(cached_id); run_end->push_back(row_base + (uint64_t)i + 1); }
    }
}

// ---- merge (merge.rs; table_function.rs:430-520) --------------------------------------------------------------------------------
static Column make_utf8(const std::string& name, const std::vector<std::string_view>& vals) {
    auto ob = std::make_shared<OwnedBuffers>();
    int64_t total = 0;
    for (auto& v : vals) total += (int64_t)v.size();
    int32_t* offs = (int32_t*)ob->alloc((vals.size() + 1) * 4);
    char* data = (char*)ob->alloc((size_t)total);
    int64_t pos = 0;
    for (size_t i = 0; i < vals.size(); i++) { offs[i] = (int32_t)pos; if (!vals[i].empty()) memcpy(data + pos, vals[i].data(), vals[i].size()); pos += (int64_t)vals[i].size(); }
    offs[vals.size()] = (int32_t)pos;
    Column c;
    c.f = Field{name, "u", false};
    c.length = (int64_t)vals.size();
    c.buffers = {nullptr, offs, data};
    c.keep = ob;
    return c;
}

static Err run_merge(brgh_session* s, const char* table, int64_t min_dist, const char* const* args, int n_args, bool cluster,
                     brgh_result* res) {
    const std::string fn = cluster ? "cluster" : "merge";
    if (min_dist < 0) return plan(fn + "() min_dist must be >= 0, got " + std::to_string(min_dist));  // table_function.rs:236-241
    // parse_merge_col_args (table_function.rs:472-520): 0 or 3 column names, optional trailing 'strict' / 'weak'
    std::vector<std::string> a;
    for (int i = 0; i < n_args; i++) a.push_back(args[i] ? args[i] : "");
    bool strict = false;
    if (!a.empty()) {
        const std::string l = lower(a.back());
        if (l == "strict") { strict = true; a.pop_back(); }
        else if (l == "weak") { a.pop_back(); }
    }
    ColTriple cols{kDefaultCols[0], kDefaultCols[1], kDefaultCols[2]};
    if (a.size() == 3) {
        for (auto& v : a) BRGH_TRY(check_name(v, "column name", fn));
        cols = ColTriple{a[0], a[1], a[2]};
    } else if (!a.empty()) {
        return plan(fn + "() expects 0 or 3 column name arguments (got " + std::to_string(a.size()) + "). Usage: " + fn +
                    "('table' [, min_dist] [, col1, col2, col3] [, 'strict'|'weak'])");
    }
    const Table* t;
    BRGH_TRY(find_table(s, table, &t));
    // StreamCollector (grouped_stream.rs:48-98): group rows by contig; groups are emitted in contig-name order
    KeyDict dict;
    std::vector<std::string_view> names;  // first-seen order
    std::vector<uint32_t> key;
    std::vector<int64_t> st64, en64;
    for (auto& b : t->batches) {
        if (b.rows == 0) continue;
        ContigArray ca;
        BRGH_TRY(contig_array(b, cols.contig, &ca));
        BRGH_TRY(resolve_positions_i64(b, cols.start, "start", &st64));
        BRGH_TRY(resolve_positions_i64(b, cols.end, "end", &en64));
        for (int64_t i = 0; i < b.rows; i++) {
            const std::string_view v = ca.value(i);
            const uint32_t id = dict.get_or_add(v);
            if (id == names.size()) names.push_back(v);
            key.push_back(id);
        }
    }
    std::vector<uint32_t> order(names.size()), rank(names.size());
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return names[x] < names[y]; });  // String::cmp = bytes
    for (uint32_t r = 0; r < order.size(); r++) rank[order[r]] = r;
    std::vector<int32_t> st(key.size()), en(key.size());
    for (size_t i = 0; i < key.size(); i++) {
        key[i] = rank[key[i]];
        if (st64[i] < INT32_MIN || st64[i] > INT32_MAX || en64[i] < INT32_MIN || en64[i] > INT32_MAX)
            return notimpl(fn + ": coordinate value " + std::to_string(st64[i] < INT32_MIN || st64[i] > INT32_MAX ? st64[i] : en64[i]) +
                           " at row " + std::to_string(i) + " does not fit the 32-bit coordinates of the GPU path");
        st[i] = (int32_t)st64[i];
        en[i] = (int32_t)en64[i];
    }
    brg_merged* m = nullptr;
    if (!key.empty()) {
        brg_ctx* ctx;
        BRGH_TRY(session_ctx(s, &ctx));
        auto f = cluster ? brg_cluster : brg_merge;
        BRGH_TRY(from_abi(f(ctx, key.data(), st.data(), en.data(), key.size(), (uint32_t)names.size(), min_dist, strict ? 1 : 0, &m)));
    }
    if (cluster) {
        // ClusterProvider's schema (cluster.rs:40-72): a 3-column input is rebuilt as (Utf8, Int64, Int64), a wider one keeps
        // all its fields (arrow take by row); then cluster / cluster_start / cluster_end Int64 NOT NULL.  Rows come out in
        // (contig, start, end, row) order, in batches of the session batch size.
        const uint64_t total = brg_merged_len(m), batch_size = 8192;
        const bool extra = t->schema.size() > 3;
        Err e = kOk;
        for (uint64_t r0 = 0; r0 < total && !e.code; r0 += batch_size) {
            const uint64_t r1 = std::min(total, r0 + batch_size);
            const uint32_t* rows = brg_merged_key(m) + r0;
            Batch ob;
            ob.rows = (int64_t)(r1 - r0);
            if (extra) {
                e = take_table(*t, rows, nullptr, ob.rows, "", -1, &ob.cols);
            } else {
                std::vector<std::string_view> cn;
                std::vector<int64_t> cs, ce;
                for (uint64_t i = r0; i < r1; i++) {
                    const uint32_t row = brg_merged_key(m)[i];
                    cn.push_back(names[order[key[row]]]);
                    cs.push_back(st64[row]);
                    ce.push_back(en64[row]);
                }
                ob.cols.push_back(make_utf8(cols.contig, cn));
                ob.cols.push_back(make_primitive<int64_t>(cols.start, "l", false, std::move(cs), nullptr, 0));
                ob.cols.push_back(make_primitive<int64_t>(cols.end, "l", false, std::move(ce), nullptr, 0));
            }
            std::vector<int64_t> cid, cst, cen;
            for (uint64_t i = r0; i < r1; i++) {
                cid.push_back(brg_merged_count(m)[i]);
                cst.push_back(brg_merged_start(m)[i]);
                cen.push_back(brg_merged_end(m)[i]);
            }
            ob.cols.push_back(make_primitive<int64_t>("cluster", "l", false, std::move(cid), nullptr, 0));
            ob.cols.push_back(make_primitive<int64_t>("cluster_start", "l", false, std::move(cst), nullptr, 0));
            ob.cols.push_back(make_primitive<int64_t>("cluster_end", "l", false, std::move(cen), nullptr, 0));
            res->batches.push_back(std::move(ob));
        }
        brg_merged_release(m);
        return e;
    }
    const uint64_t total = brg_merged_len(m);
    const uint64_t batch_size = 8192;  // the session's batch size: MergeStream flushes its builders at this many rows
    for (uint64_t r0 = 0; r0 < total; r0 += batch_size) {
        const uint64_t r1 = std::min(total, r0 + batch_size);
        std::vector<std::string_view> cn;
        std::vector<int64_t> cs, ce, cc;
        for (uint64_t i = r0; i < r1; i++) {
            cn.push_back(names[order[brg_merged_key(m)[i]]]);
            cs.push_back(brg_merged_start(m)[i]);
            ce.push_back(brg_merged_end(m)[i]);
            cc.push_back(brg_merged_count(m)[i]);
        }
        Batch ob;
        ob.rows = (int64_t)(r1 - r0);
        ob.cols.push_back(make_utf8(cols.contig, cn));
        ob.cols.push_back(make_primitive<int64_t>(cols.start, "l", false, std::move(cs), nullptr, 0));
        ob.cols.push_back(make_primitive<int64_t>(cols.end, "l", false, std::move(ce), nullptr, 0));
        ob.cols.push_back(make_primitive<int64_t>("n_intervals", "l", false, std::move(cc), nullptr, 0));
        res->batches.push_back(std::move(ob));
    }
    brg_merged_release(m);
    return kOk;
}

// ---- subtract (subtract.rs; table_function.rs) -----------------------------------------------------------------------------------
static Err run_subtract(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args, brgh_result* res) {
    UdtfArgs ua;
    BRGH_TRY(parse_col_args(args, n_args, "subtract", &ua));
    const Table *lt, *rt;
    BRGH_TRY(find_table(s, left, &lt));
    BRGH_TRY(find_table(s, right, &rt));
    // right side: merged intervals per contig on the device (the sweep walks them, see brg_subtract)
    KeyDict rdict;
    IndexGuard ig;
    BRGH_TRY(build_from_table(s, *rt, ua.right, BRG_WANT_COVERAGE, &rdict, &ig));
    // left side: rows in (contig name, start, end, row) order (FullBatchCollector::take_groups, grouped_stream.rs:228-237)
    KeyDict ldict;
    std::vector<std::string_view> names;
    std::vector<uint32_t> sort_key, probe_key;
    std::vector<int64_t> st64, en64;
    for (auto& b : lt->batches) {
        if (b.rows == 0) continue;
        ContigArray ca;
        BRGH_TRY(contig_array(b, ua.left.contig, &ca));
        BRGH_TRY(resolve_positions_i64(b, ua.left.start, "start", &st64));
        BRGH_TRY(resolve_positions_i64(b, ua.left.end, "end", &en64));
        for (int64_t i = 0; i < b.rows; i++) {
            const std::string_view v = ca.value(i);
            const uint32_t id = ldict.get_or_add(v);
            if (id == names.size()) names.push_back(v);
            sort_key.push_back(id);
            probe_key.push_back(rdict.lookup(v));
        }
    }
    const uint64_t n = sort_key.size();
    std::vector<uint32_t> name_order(names.size()), rank(names.size());
    for (uint32_t i = 0; i < name_order.size(); i++) name_order[i] = i;
    std::sort(name_order.begin(), name_order.end(), [&](uint32_t x, uint32_t y) { return names[x] < names[y]; });
    for (uint32_t r = 0; r < name_order.size(); r++) rank[name_order[r]] = r;
    std::vector<int32_t> st(n), en(n);
    for (uint64_t i = 0; i < n; i++) {
        sort_key[i] = rank[sort_key[i]];
        if (st64[i] < INT32_MIN || st64[i] > INT32_MAX || en64[i] < INT32_MIN || en64[i] > INT32_MAX)
            return notimpl("subtract: coordinate value " + std::to_string(st64[i] < INT32_MIN || st64[i] > INT32_MAX ? st64[i] : en64[i]) +
                           " at row " + std::to_string(i) + " does not fit the 32-bit coordinates of the GPU path");
        st[i] = (int32_t)st64[i];
        en[i] = (int32_t)en64[i];
    }
    if (n == 0) return kOk;
    std::vector<uint32_t> order(n);
    BRGH_TRY(from_abi(brg_sort_rows(s->ctx, sort_key.data(), st.data(), en.data(), n, std::max<uint32_t>(1, (uint32_t)names.size()), order.data())));
    std::vector<uint32_t> pk(n);
    std::vector<int32_t> ps(n), pe(n);
    for (uint64_t i = 0; i < n; i++) { pk[i] = probe_key[order[i]]; ps[i] = st[order[i]]; pe[i] = en[order[i]]; }
    brg_rows* rows = nullptr;
    BRGH_TRY(from_abi(brg_subtract(s->ctx, ig.ix, pk.data(), ps.data(), pe.data(), n, ua.strict ? 1 : 0, &rows)));
    // output: a 3-column left table is rebuilt as (Utf8, Int64, Int64); a wider one keeps its columns, with the start / end
    // columns replaced by the Int64 pieces (subtract.rs:48-75)
    const uint64_t total = brg_rows_len(rows), batch_size = 8192;
    const bool extra = lt->schema.size() > 3;
    Err e = kOk;
    for (uint64_t r0 = 0; r0 < total && !e.code; r0 += batch_size) {
        const uint64_t r1 = std::min(total, r0 + batch_size);
        std::vector<uint32_t> src(r1 - r0);
        std::vector<int64_t> cs, ce;
        for (uint64_t i = r0; i < r1; i++) {
            src[i - r0] = order[brg_rows_probe_idx(rows)[i]];
            cs.push_back(brg_rows_piece_start(rows)[i]);
            ce.push_back(brg_rows_piece_end(rows)[i]);
        }
        Batch ob;
        ob.rows = (int64_t)(r1 - r0);
        if (extra) {
            e = take_table(*lt, src.data(), nullptr, ob.rows, "", -1, &ob.cols);
            for (auto& c : ob.cols) {
                if (c.f.name == ua.left.start) c = make_primitive<int64_t>(c.f.name, "l", c.f.nullable, std::move(cs), nullptr, 0);
                else if (c.f.name == ua.left.end) c = make_primitive<int64_t>(c.f.name, "l", c.f.nullable, std::move(ce), nullptr, 0);
            }
        } else {
            std::vector<std::string_view> cn;
            for (uint32_t row : src) cn.push_back(names[name_order[sort_key[row]]]);
            ob.cols.push_back(make_utf8(ua.left.contig, cn));
            ob.cols.push_back(make_primitive<int64_t>(ua.left.start, "l", false, std::move(cs), nullptr, 0));
            ob.cols.push_back(make_primitive<int64_t>(ua.left.end, "l", false, std::move(ce), nullptr, 0));
        }
        res->batches.push_back(std::move(ob));
    }
    brg_rows_release(rows);
    return e;
}

// ---- complement (complement.rs; table_function.rs:602-767) -------------------------------------------------------------------
static Err run_complement(brgh_session* s, const char* table, const char* const* args, int n_args, brgh_result* res) {
    const std::string fn = "complement";
    // argument grammar of ComplementTableFunction::call (table_function.rs:615-767): after the table name, an optional trailing
    // 'strict' / 'weak', then 0 / 1 (view table) / 3 (columns) / 4 (view + columns) / 7 (view + columns + view columns) strings
    std::vector<std::string> a;
    for (int i = 0; i < n_args; i++) a.push_back(args[i] ? args[i] : "");
    bool strict = false;
    if (!a.empty()) {
        const std::string l = lower(a.back());
        if (l == "strict") { strict = true; a.pop_back(); }
        else if (l == "weak") { a.pop_back(); }
    }
    ColTriple cols{kDefaultCols[0], kDefaultCols[1], kDefaultCols[2]};
    ColTriple vcols = cols;
    std::string view_table;
    bool has_view = false;
    switch (a.size()) {
        case 0: break;
        case 1: view_table = a[0]; has_view = true; break;
        case 3: cols = vcols = ColTriple{a[0], a[1], a[2]}; break;
        case 4: view_table = a[0]; has_view = true; cols = vcols = ColTriple{a[1], a[2], a[3]}; break;
        case 7: view_table = a[0]; has_view = true; cols = ColTriple{a[1], a[2], a[3]}; vcols = ColTriple{a[4], a[5], a[6]}; break;
        default:
            return plan(fn + "() unexpected number of arguments (" + std::to_string(a.size()) + " extra args after table name). Usage: " + fn +
                        "('table' [, 'view_table'] [, col1, col2, col3 [, vcol1, vcol2, vcol3]] [, 'strict'|'weak'])");
    }
    const Table *t, *vt = nullptr;
    BRGH_TRY(find_table(s, table, &t));
    if (has_view) BRGH_TRY(find_table(s, view_table.c_str(), &vt));
    // input rows grouped by contig (StreamCollector, grouped_stream.rs:48-98); groups are walked in contig-name order
    KeyDict dict;
    std::vector<std::string_view> names;
    std::vector<uint32_t> key;
    std::vector<int64_t> st64, en64;
    for (auto& b : t->batches) {
        if (b.rows == 0) continue;
        ContigArray ca;
        BRGH_TRY(contig_array(b, cols.contig, &ca));
        BRGH_TRY(resolve_positions_i64(b, cols.start, "start", &st64));
        BRGH_TRY(resolve_positions_i64(b, cols.end, "end", &en64));
        for (int64_t i = 0; i < b.rows; i++) {
            const std::string_view v = ca.value(i);
            const uint32_t id = dict.get_or_add(v);
            if (id == names.size()) names.push_back(v);
            key.push_back(id);
        }
    }
    std::vector<uint32_t> order(names.size()), rank(names.size());
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return names[x] < names[y]; });
    for (uint32_t r = 0; r < order.size(); r++) rank[order[r]] = r;
    std::vector<int32_t> st(key.size()), en(key.size());
    for (size_t i = 0; i < key.size(); i++) {
        key[i] = rank[key[i]];
        if (st64[i] < INT32_MIN || st64[i] > INT32_MAX || en64[i] < INT32_MIN || en64[i] > INT32_MAX)
            return notimpl(fn + ": coordinate value " + std::to_string(st64[i] < INT32_MIN || st64[i] > INT32_MAX ? st64[i] : en64[i]) +
                           " at row " + std::to_string(i) + " does not fit the 32-bit coordinates of the GPU path");
        st[i] = (int32_t)st64[i];
        en[i] = (int32_t)en64[i];
    }
    // view rows per contig, each contig's rows in (start, end) order (take_groups_as_map after sort_unstable)
    KeyDict vdict;
    std::vector<std::string_view> vnames;
    std::vector<std::vector<std::pair<int64_t, int64_t>>> vgroups;
    if (vt) {
        std::vector<int64_t> vs64, ve64;
        for (auto& b : vt->batches) {
            if (b.rows == 0) continue;
            ContigArray ca;
            BRGH_TRY(contig_array(b, vcols.contig, &ca));
            vs64.clear(); ve64.clear();
            BRGH_TRY(resolve_positions_i64(b, vcols.start, "start", &vs64));
            BRGH_TRY(resolve_positions_i64(b, vcols.end, "end", &ve64));
            for (int64_t i = 0; i < b.rows; i++) {
                const std::string_view v = ca.value(i);
                const uint32_t id = vdict.get_or_add(v);
                if (id == vnames.size()) { vnames.push_back(v); vgroups.emplace_back(); }
                vgroups[id].emplace_back(vs64[i], ve64[i]);
            }
        }
        for (auto& g : vgroups) std::sort(g.begin(), g.end());
    }
    // emission order (complement.rs:396-457): input contigs by name, each with its views or the implicit (0, i64::MAX);
    // then the view-only contigs by name, whole
    std::vector<uint32_t> vkey;
    std::vector<int64_t> vstart, vend;
    std::vector<std::string_view> vcontig;
    for (uint32_t r = 0; r < order.size(); r++) {
        const std::string_view nm = names[order[r]];
        const uint32_t vid = vdict.lookup(nm);
        if (vid == BRG_KEY_ABSENT) {
            vkey.push_back(r); vstart.push_back(0); vend.push_back(INT64_MAX); vcontig.push_back(nm);
        } else {
            for (auto& pr : vgroups[vid]) { vkey.push_back(r); vstart.push_back(pr.first); vend.push_back(pr.second); vcontig.push_back(nm); }
        }
    }
    std::vector<uint32_t> vorder(vnames.size());
    for (uint32_t i = 0; i < vorder.size(); i++) vorder[i] = i;
    std::sort(vorder.begin(), vorder.end(), [&](uint32_t x, uint32_t y) { return vnames[x] < vnames[y]; });
    for (uint32_t vi : vorder) {
        if (dict.lookup(vnames[vi]) != BRG_KEY_ABSENT) continue;
        for (auto& pr : vgroups[vi]) { vkey.push_back(BRG_KEY_ABSENT); vstart.push_back(pr.first); vend.push_back(pr.second); vcontig.push_back(vnames[vi]); }
    }
    if (vkey.empty()) return kOk;
    brg_ctx* ctx;
    BRGH_TRY(session_ctx(s, &ctx));
    brg_gaps* g = nullptr;
    BRGH_TRY(from_abi(brg_complement(ctx, key.data(), st.data(), en.data(), key.size(), std::max<uint32_t>(1, (uint32_t)names.size()),
                                     vkey.data(), vstart.data(), vend.data(), vkey.size(), strict ? 1 : 0, &g)));
    // ComplementProvider's schema (complement.rs:46-50): (contig Utf8, start Int64, end Int64) NOT NULL, named after the input columns
    const uint64_t total = brg_gaps_len(g), batch_size = 8192;
    for (uint64_t r0 = 0; r0 < total; r0 += batch_size) {
        const uint64_t r1 = std::min(total, r0 + batch_size);
        std::vector<std::string_view> cn;
        std::vector<int64_t> cs, ce;
        for (uint64_t i = r0; i < r1; i++) {
            cn.push_back(vcontig[brg_gaps_view(g)[i]]);
            cs.push_back(brg_gaps_start(g)[i]);
            ce.push_back(brg_gaps_end(g)[i]);
        }
        Batch ob;
        ob.rows = (int64_t)(r1 - r0);
        ob.cols.push_back(make_utf8(cols.contig, cn));
        ob.cols.push_back(make_primitive<int64_t>(cols.start, "l", false, std::move(cs), nullptr, 0));
        ob.cols.push_back(make_primitive<int64_t>(cols.end, "l", false, std::move(ce), nullptr, 0));
        res->batches.push_back(std::move(ob));
    }
    brg_gaps_release(g);
    return kOk;
}

// ---- count_overlaps / coverage (count_overlaps.rs:107-169; interval_tree.rs:155-280) ----------------------------------------
static Err run_count(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args, bool coverage,
                     brgh_result* res) {
    const std::string fn = coverage ? "coverage" : "count_overlaps";
    UdtfArgs ua;
    BRGH_TRY(parse_col_args(args, n_args, fn, &ua));
    const Table *lt, *rt;
    BRGH_TRY(find_table(s, left, &lt));
    BRGH_TRY(find_table(s, right, &rt));
    KeyDict dict;
    IndexGuard ig;
    BRGH_TRY(build_from_table(s, *lt, ua.left, coverage ? BRG_WANT_COVERAGE : BRG_WANT_COUNT, &dict, &ig));
    // probe batches are coalesced up to coalesce_rows per device round trip; the output keeps the input batching
    size_t b0 = 0;
    while (b0 < rt->batches.size()) {
        // the position value buffers go to the device in their Arrow width (Int32 / Int64 / UInt32 / UInt64) and are narrowed
        // there (brg_*_arrow): no per-row host loop
        std::vector<uint8_t> sraw, eraw;
        int s_type = BRG_POS_INT32, e_type = BRG_POS_INT32;
        std::vector<uint32_t> run_key;
        std::vector<uint64_t> run_end, batch_first_row;
        size_t b1 = b0;
        uint64_t rows = 0;
        while (b1 < rt->batches.size() && (b1 == b0 || rows + (uint64_t)rt->batches[b1].rows <= s->coalesce_rows)) {
            const Batch& b = rt->batches[b1];
            ContigArray ca;
            BRGH_TRY(contig_array(b, ua.right.contig, &ca));
            BRGH_TRY(position_bytes(b, ua.right.start, "start", &sraw, &s_type));
            BRGH_TRY(position_bytes(b, ua.right.end, "end", &eraw, &e_type));
            probe_key_runs(ca, b.rows, dict, false, rows, &run_key, &run_end);
            batch_first_row.push_back(rows);
            rows += (uint64_t)b.rows;
            b1++;
        }
        std::vector<int64_t> out(rows);
        if (rows) {
            auto f = coverage ? brg_coverage_arrow : brg_count_overlaps_arrow;
            uint64_t bad[2];
            const int rc = f(s->ctx, ig.ix, run_key.data(), run_end.data(), run_key.size(), sraw.data(), s_type, eraw.data(), e_type, rows,
                             ua.strict ? 1 : 0, out.data(), bad);
            if (rc == BRG_ERR_OVERFLOW && (bad[0] != ~0ull || bad[1] != ~0ull)) return ingest_overflow(bad, batch_first_row, sraw, s_type, eraw, e_type);
            BRGH_TRY(from_abi(rc));
        }
        uint64_t pos = 0;
        for (size_t bi = b0; bi < b1; bi++) {
            Batch ob = rt->batches[bi];  // the probe batch's columns, zero-copy (interval_tree.rs:269-273)
            std::vector<int64_t> col(out.begin() + pos, out.begin() + pos + ob.rows);
            pos += (uint64_t)ob.rows;
            ob.cols.push_back(make_primitive<int64_t>(coverage ? "coverage" : "count", "l", false, std::move(col), nullptr, 0));
            res->batches.push_back(std::move(ob));
        }
        b0 = b1;
    }
    return kOk;
}

// ---- nearest (nearest.rs:127-547) -------------------------------------------------------------------------------------------------
static Err run_nearest(brgh_session* s, const char* left, const char* right, int64_t k, bool include_overlaps,
                       bool compute_distance, const char* const* args, int n_args, brgh_result* res) {
    if (k < 1) return plan("nearest() k must be >= 1");  // table_function.rs:200-205
    UdtfArgs ua;
    BRGH_TRY(parse_col_args(args, n_args, "nearest", &ua));
    const Table *lt, *rt;
    BRGH_TRY(find_table(s, left, &lt));
    BRGH_TRY(find_table(s, right, &rt));
    KeyDict dict;
    IndexGuard ig;
    BRGH_TRY(build_from_table(s, *lt, ua.left, BRG_WANT_NEAREST, &dict, &ig));
    for (auto& b : rt->batches) {
        ContigArray ca;
        std::vector<int32_t> st, en;
        std::vector<uint8_t> sraw, eraw;  // k == 1: value buffers in their Arrow width, narrowed on the device
        int s_type = BRG_POS_INT32, e_type = BRG_POS_INT32;
        BRGH_TRY(contig_array(b, ua.right.contig, &ca));
        if (k == 1) {
            BRGH_TRY(position_bytes(b, ua.right.start, "start", &sraw, &s_type));
            BRGH_TRY(position_bytes(b, ua.right.end, "end", &eraw, &e_type));
        } else {
            BRGH_TRY(resolve_positions(b, ua.right.start, "start", &st));
            BRGH_TRY(resolve_positions(b, ua.right.end, "end", &en));
        }
        std::vector<uint32_t> run_key;
        std::vector<uint64_t> run_end;
        probe_key_runs(ca, b.rows, dict, true, 0, &run_key, &run_end);
        const uint64_t n = (uint64_t)b.rows;
        std::vector<uint32_t> left_idx, right_idx;
        std::vector<uint8_t> bits;
        std::vector<int64_t> dist;
        uint64_t m = 0;
        if (k == 1) {
            left_idx.resize(n);
            bits.assign(((n + 31) / 32) * 4, 0);
            if (compute_distance) dist.resize(n);
            if (n) {
                uint64_t bad[2];
                const int rc = brg_nearest_k1_arrow(s->ctx, ig.ix, run_key.data(), run_end.data(), run_key.size(), sraw.data(), s_type,
                                                    eraw.data(), e_type, n, include_overlaps, ua.strict, left_idx.data(), bits.data(),
                                                    compute_distance ? dist.data() : nullptr, bad);
                if (rc == BRG_ERR_OVERFLOW && (bad[0] != ~0ull || bad[1] != ~0ull))
                    return ingest_overflow(bad, std::vector<uint64_t>{0}, sraw, s_type, eraw, e_type);
                BRGH_TRY(from_abi(rc));
            }
            right_idx.resize(n);
            for (uint64_t i = 0; i < n; i++) right_idx[i] = (uint32_t)i;
            m = n;
        } else if (n) {
            // k > 1 goes through the per-row key column form
            std::vector<uint32_t> key(n);
            uint64_t r0 = 0;
            for (size_t r = 0; r < run_key.size(); r++) { for (; r0 < run_end[r]; r0++) key[r0] = run_key[r]; }
            brg_rows* rows = nullptr;
            BRGH_TRY(from_abi(brg_nearest(s->ctx, ig.ix, key.data(), st.data(), en.data(), n, (uint32_t)std::min<int64_t>(k, 0xFFFFFFFFll),
                                          include_overlaps, compute_distance, ua.strict, &rows)));
            m = brg_rows_len(rows);
            left_idx.assign(brg_rows_build_idx(rows), brg_rows_build_idx(rows) + m);
            right_idx.assign(brg_rows_probe_idx(rows), brg_rows_probe_idx(rows) + m);
            bits.assign(brg_rows_valid_bits(rows), brg_rows_valid_bits(rows) + ((m + 31) / 32) * 4);
            if (compute_distance) dist.assign(brg_rows_distance(rows), brg_rows_distance(rows) + m);
            brg_rows_release(rows);
        }
        int64_t nulls = 0;
        for (uint64_t i = 0; i < m; i++) nulls += ((bits[i >> 3] >> (i & 7)) & 1) ? 0 : 1;
        Batch ob;
        ob.rows = (int64_t)m;
        // take(left columns, left_idx with nulls) ++ take(right columns, right_idx) ++ distance (nearest.rs:456-482)
        BRGH_TRY(take_table(*lt, left_idx.data(), nulls ? bits.data() : nullptr, (int64_t)m, "left_", 1, &ob.cols));
        Table rb;
        rb.schema = rt->schema;
        rb.batches.push_back(b);
        BRGH_TRY(take_table(rb, right_idx.data(), nullptr, (int64_t)m, "right_", 1, &ob.cols));
        if (compute_distance) {
            dist.resize(m);
            ob.cols.push_back(make_primitive<int64_t>("distance", "l", true, std::move(dist), bits.data(), nulls));
        }
        res->batches.push_back(std::move(ob));
    }
    return kOk;
}

// ---- IntervalJoinExec (interval_join.rs) ---------------------------------------------------------------------------------------------
// byte-string encoding of the `on` tuple of one row: per column a NULL marker or the value bytes, length-prefixed
// (the reference hashes the tuple and so lets NULL keys meet each other, interval_join.rs:917,1128)
static Err append_key_bytes(const Column& c, int64_t i, std::string* k) {
    if (c.is_null(i)) { k->push_back('\0'); return kOk; }
    k->push_back('\1');
    const std::string& f = c.f.format;
    const int w = fixed_width(f);
    if (w) { k->append((const char*)c.buffers[1] + (c.offset + i) * w, (size_t)w); return kOk; }
    if (f == "u" || f == "U" || f == "vu") {
        ContigArray ca{&c, f == "u" ? 0 : (f == "U" ? 1 : 2)};
        const std::string_view v = ca.value(i);
        const uint32_t len = (uint32_t)v.size();
        k->append((const char*)&len, 4);
        k->append(v.data(), v.size());
        return kOk;
    }
    return notimpl("join key column '" + c.f.name + "' of type " + type_name(f) + " is not supported by the host layer");
}
static Err key_columns(const Batch& b, const char* const* names, int n, std::vector<const Column*>* out) {
    for (int i = 0; i < n; i++) {
        const Column* c = b.find(names[i] ? names[i] : "");
        if (!c) return plan(std::string("join key column '") + (names[i] ? names[i] : "") + "' not found in batch with columns: " + b.names());
        out->push_back(c);
    }
    return kOk;
}

static uint64_t max_output_batch_size() {  // interval_join.rs:543-548
    const char* e = getenv("BIO_MAX_OUTPUT_BATCH_SIZE");
    if (!e) return 100000;
    char* end = nullptr;
    const unsigned long long v = strtoull(e, &end, 10);
    return (end && *end == 0 && v > 0) ? v : 100000;
}

static Err run_join(brgh_session* s, const Table& lt, const Table& rt, const char* const* on_left, const char* const* on_right,
                    int n_on, const std::string& ls, const std::string& le, const std::string& rs, const std::string& re,
                    bool lt_strict, bool gt_strict, int join_type, const std::string& left_prefix, const std::string& right_prefix,
                    bool emit_left, bool emit_right, brgh_result* res) {
    // ---- build side = SQL-left input, collected once (interval_join.rs:584-668) ------------------------------------------
    KeyDict dict;
    std::vector<uint32_t> key;
    std::vector<int32_t> bst, ben;
    for (auto& b : lt.batches) {
        std::vector<const Column*> kc;
        BRGH_TRY(key_columns(b, on_left, n_on, &kc));
        BRGH_TRY(resolve_positions(b, ls, "start", &bst, true));
        BRGH_TRY(resolve_positions(b, le, "end", &ben, true));
        std::string kb;
        for (int64_t i = 0; i < b.rows; i++) {
            kb.clear();
            for (auto* c : kc) BRGH_TRY(append_key_bytes(*c, i, &kb));
            key.push_back(dict.get_or_add(kb));
        }
    }
    // `l.end > r.start` compares against l.end - 1 (intervals.rs:85-115: minus_one on the side of the strict comparison)
    if (gt_strict) for (auto& v : ben) v = (int32_t)((uint32_t)v - 1u);
    // one-row-per-probe modes (interval_join.rs:1627-1636): CoitreesNearest = nearest_one(start, end, true);
    // CoitreesCountOverlaps = the first overlapping interval or NULL.  The reference takes the first interval its tree
    // visits (order unspecified); here it is the first in (start, end, row id) order.
    const bool nearest_mode = s->algorithm == CoitreesNearest || s->algorithm == CoitreesCountOverlaps;
    const int nearest_overlaps = s->algorithm == CoitreesCountOverlaps ? BRG_NEAREST_OVERLAP_ONLY : 1;
    brg_ctx* ctx;
    BRGH_TRY(session_ctx(s, &ctx));
    IndexGuard ig;
    const int mode = join_type == BRGH_JOIN_RIGHT_SEMI ? BRG_PAIRS_SEMI : (join_type == BRGH_JOIN_RIGHT_ANTI ? BRG_PAIRS_ANTI : BRG_PAIRS_INNER);
    // Existence joins ask IntervalJoinAlgorithm::get whether ANY interval is visited (interval_join.rs:1433-1463): under
    // CoitreesCountOverlaps that is the plain tree query (:854-863) -> the pair structures; under CoitreesNearest it is
    // nearest_one(start, end, true) (:864-870) -> every probe row of a key with a non-empty index "matches".
    const bool exist_by_nearest = mode != BRG_PAIRS_INNER && s->algorithm == CoitreesNearest;
    const bool want_nearest = (nearest_mode && mode == BRG_PAIRS_INNER) || exist_by_nearest;
    BRGH_TRY(from_abi(brg_build(ctx, key.data(), bst.data(), ben.data(), key.size(), std::max<uint32_t>(1, (uint32_t)dict.ids.size()),
                                want_nearest ? BRG_WANT_NEAREST : BRG_WANT_PAIRS, &ig.ix)));
    const uint64_t cap = s->low_memory ? max_output_batch_size() : 0;
    // ---- probe side, batch by batch (interval_join.rs:1418-1676) ----------------------------------------------------------
    for (auto& b : rt.batches) {
        std::vector<const Column*> kc;
        BRGH_TRY(key_columns(b, on_right, n_on, &kc));
        std::vector<int32_t> pst, pen;
        BRGH_TRY(resolve_positions(b, rs, "start", &pst, true));
        BRGH_TRY(resolve_positions(b, re, "end", &pen, true));
        if (lt_strict) for (auto& v : pen) v = (int32_t)((uint32_t)v - 1u);  // `l.start < r.end` -> r.end - 1
        std::vector<uint32_t> pk((size_t)b.rows);
        std::string kb;
        for (int64_t i = 0; i < b.rows; i++) {
            kb.clear();
            for (auto* c : kc) BRGH_TRY(append_key_bytes(*c, i, &kb));
            pk[(size_t)i] = dict.lookup(kb);
        }
        Table rb;
        rb.schema = rt.schema;
        rb.batches.push_back(b);
        const uint64_t n = (uint64_t)b.rows;
        if (nearest_mode && mode == BRG_PAIRS_INNER) {
            // one row per probe row: nearest_one(start, end, true), NULL left when the key has no index (interval_join.rs:864-870)
            std::vector<uint32_t> li(n);
            std::vector<uint8_t> bits(((n + 31) / 32) * 4, 0);
            if (n) BRGH_TRY(from_abi(brg_nearest_k1(ctx, ig.ix, pk.data(), pst.data(), pen.data(), n, nearest_overlaps, 0, li.data(), bits.data(), nullptr)));
            int64_t nulls = 0;
            for (uint64_t i = 0; i < n; i++) nulls += ((bits[i >> 3] >> (i & 7)) & 1) ? 0 : 1;
            std::vector<uint32_t> ri(n);
            for (uint64_t i = 0; i < n; i++) ri[i] = (uint32_t)i;
            Batch ob;
            ob.rows = (int64_t)n;
            if (emit_left) BRGH_TRY(take_table(lt, li.data(), nulls ? bits.data() : nullptr, (int64_t)n, left_prefix, -1, &ob.cols));
            if (emit_right) BRGH_TRY(take_table(rb, ri.data(), nullptr, (int64_t)n, right_prefix, -1, &ob.cols));
            res->batches.push_back(std::move(ob));
            continue;
        }
        if (exist_by_nearest) {
            std::vector<uint32_t> li(n);
            std::vector<uint8_t> bits(((n + 31) / 32) * 4, 0);
            if (n) BRGH_TRY(from_abi(brg_nearest_k1(ctx, ig.ix, pk.data(), pst.data(), pen.data(), n, 1, 0, li.data(), bits.data(), nullptr)));
            std::vector<uint32_t> pi;
            for (uint64_t i = 0; i < n; i++) {
                const bool has_match = (bits[i >> 3] >> (i & 7)) & 1;
                if (has_match == (mode == BRG_PAIRS_SEMI)) pi.push_back((uint32_t)i);  // should_emit_probe_row
            }
            if (pi.empty()) continue;  // (the reference emits no batch for a probe batch without surviving rows)
            Batch ob;
            ob.rows = (int64_t)pi.size();
            BRGH_TRY(take_table(rb, pi.data(), nullptr, ob.rows, right_prefix, -1, &ob.cols));
            res->batches.push_back(std::move(ob));
            continue;
        }
        uint64_t done = 0;
        do {  // low-memory mode re-enters with the rows not consumed yet (interval_join.rs:1466-1573)
            brg_rows* rows = nullptr;
            uint64_t consumed = 0;
            BRGH_TRY(from_abi(brg_overlap_pairs(ctx, ig.ix, pk.data() + done, pst.data() + done, pen.data() + done, n - done, mode, cap,
                                                &rows, &consumed)));
            const uint64_t m = brg_rows_len(rows);
            std::vector<uint32_t> pi(brg_rows_probe_idx(rows), brg_rows_probe_idx(rows) + m);
            for (auto& v : pi) v += (uint32_t)done;
            Batch ob;
            ob.rows = (int64_t)m;
            Err e = kOk;
            if (mode == BRG_PAIRS_INNER) {
                if (emit_left) e = take_table(lt, brg_rows_build_idx(rows), nullptr, (int64_t)m, left_prefix, -1, &ob.cols);
                if (!e.code && emit_right) e = take_table(rb, pi.data(), nullptr, (int64_t)m, right_prefix, -1, &ob.cols);
            } else {
                e = take_table(rb, pi.data(), nullptr, (int64_t)m, right_prefix, -1, &ob.cols);  // RightSemi/RightAnti: probe columns only
            }
            brg_rows_release(rows);
            BRGH_TRY(e);
            res->batches.push_back(std::move(ob));
            done += std::max<uint64_t>(consumed, n == done ? 0 : 1);
        } while (done < n);
    }
    return kOk;
}

// ---- the range-join filter parser (physical_planner/intervals.rs:30-218) -----------------------------------------------------
// The reference walks DataFusion's JoinFilter: an AND of exactly two comparisons, each between an expression over ONE column of
// the left input and an expression over ONE column of the right input.  It assigns the four bounds by the side of the operands
// and the direction of the comparison, and rewrites a strict comparison by subtracting one from the END it involves
// (`minus_one`, intervals.rs:63-65).  DataFusion itself is out of scope here, so the filter arrives as the SQL text of the ON
// clause's residual; column references are `alias.column` (optionally quoted with " or `, optionally inside CAST(.. AS type) as
// the reference's own OverlapProvider writes them, overlap.rs:171-181).
struct FilterOperand {
    int side = -1;  // 0 = left, 1 = right
    std::string column;
};
struct FilterCmp {
    FilterOperand a, b;
    std::string op;
};
struct FilterLexer {
    const std::string& s;
    size_t i = 0;
    explicit FilterLexer(const std::string& t) : s(t) {}
    void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
    bool eat_word(const char* w) {  // case-insensitive keyword followed by a non-identifier character
        ws();
        size_t n = strlen(w);
        if (i + n > s.size()) return false;
        for (size_t k = 0; k < n; k++) if (tolower((unsigned char)s[i + k]) != tolower((unsigned char)w[k])) return false;
        if (i + n < s.size() && (isalnum((unsigned char)s[i + n]) || s[i + n] == '_')) return false;
        i += n;
        return true;
    }
    bool ident(std::string* out) {
        ws();
        out->clear();
        if (i < s.size() && (s[i] == '"' || s[i] == '`')) {
            const char q = s[i++];
            while (i < s.size() && s[i] != q) out->push_back(s[i++]);
            if (i >= s.size()) return false;
            i++;
            return true;
        }
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_')) out->push_back(s[i++]);
        return !out->empty();
    }
};
// one side of a comparison: [CAST(] alias.column [AS type)] possibly parenthesised.  Anything that combines two columns (the
// reference panics with "complex sub queries are not supported", intervals.rs:47-52) or no column at all is refused.
static Err parse_filter_operand(FilterLexer& lx, const std::string& la, const std::string& ra, FilterOperand* out) {
    if (lx.eat('(')) {
        BRGH_TRY(parse_filter_operand(lx, la, ra, out));
        lx.ws();
        if (lx.i < lx.s.size() && strchr("+-*/", lx.s[lx.i])) return internal("complex sub queries are not supported " + lx.s);
        if (!lx.eat(')')) return internal("filter expression: missing ')' in " + lx.s);
        return kOk;
    }
    if (lx.eat_word("CAST")) {
        if (!lx.eat('(')) return internal("filter expression: CAST without '(' in " + lx.s);
        BRGH_TRY(parse_filter_operand(lx, la, ra, out));
        if (!lx.eat_word("AS")) return internal("filter expression: CAST without AS in " + lx.s);
        std::string ty;
        if (!lx.ident(&ty)) return internal("filter expression: CAST without a type in " + lx.s);
        if (!lx.eat(')')) return internal("filter expression: missing ')' after CAST in " + lx.s);
        return kOk;
    }
    std::string alias, col;
    if (!lx.ident(&alias) || !lx.eat('.') || !lx.ident(&col)) return internal("filter expression: expected alias.column in " + lx.s);
    if (lower(alias) == lower(la)) out->side = 0;
    else if (lower(alias) == lower(ra)) out->side = 1;
    else return internal("filter expression: unknown table alias '" + alias + "'");
    out->column = col;
    return kOk;
}
static Err parse_filter_cmp(FilterLexer& lx, const std::string& la, const std::string& ra, FilterCmp* out) {
    const bool paren = [&] { const size_t save = lx.i; if (lx.eat('(')) { FilterOperand probe; FilterLexer t(lx.s); t.i = lx.i; Err e = parse_filter_operand(t, la, ra, &probe);
                                 t.ws(); if (!e.code && t.i < t.s.size() && strchr("<>=!", t.s[t.i])) return true; } lx.i = save; return false; }();
    BRGH_TRY(parse_filter_operand(lx, la, ra, &out->a));
    lx.ws();
    out->op.clear();
    while (lx.i < lx.s.size() && strchr("<>=!", lx.s[lx.i])) out->op.push_back(lx.s[lx.i++]);
    if (out->op.empty()) {
        if (lx.i < lx.s.size() && strchr("+-*/", lx.s[lx.i])) return internal("complex sub queries are not supported " + lx.s);
        return internal("filter expression is not a BinaryExpr");
    }
    BRGH_TRY(parse_filter_operand(lx, la, ra, &out->b));
    lx.ws();
    if (lx.i < lx.s.size() && strchr("+-*/", lx.s[lx.i])) return internal("complex sub queries are not supported " + lx.s);
    if (paren && !lx.eat(')')) return internal("filter expression: missing ')' in " + lx.s);
    return kOk;
}
struct ColIntervalsOut {
    std::string ls, le, rs, re;
    bool have_ls = false, have_le = false, have_rs = false, have_re = false;
    bool le_minus_one = false, re_minus_one = false;
};
// parse_condition (intervals.rs:67-124): which bound each operand is, by side and comparison direction
static Err apply_filter_cmp(const FilterCmp& c, ColIntervalsOut* o) {
    const bool is_lt = c.op == "<", is_lteq = c.op == "<=", is_gteq = c.op == ">=", is_gt = c.op == ">";
    if (!(is_lt || is_lteq || is_gteq || is_gt)) return internal("Unsupported operator: " + c.op);
    auto set = [&](bool* have, std::string* dst, const std::string& v, const char* what) -> Err {
        if (*have) return internal(std::string(what) + " already set");
        *have = true; *dst = v;
        return kOk;
    };
    if (c.a.side == 1 && (is_lteq || is_lt)) {          // rs </<= le
        if (c.b.side != 0) return internal("couldn't parse as rs </<= le");
        BRGH_TRY(set(&o->have_rs, &o->rs, c.a.column, "rs"));
        BRGH_TRY(set(&o->have_le, &o->le, c.b.column, "le"));
        o->le_minus_one = is_lt;
    } else if (c.a.side == 0 && (is_lteq || is_lt)) {   // ls </<= re
        if (c.b.side != 1) return internal("couldn't parse as ls </<= re");
        BRGH_TRY(set(&o->have_re, &o->re, c.b.column, "re"));
        BRGH_TRY(set(&o->have_ls, &o->ls, c.a.column, "ls"));
        o->re_minus_one = is_lt;
    } else if (c.a.side == 1 && (is_gteq || is_gt)) {   // re >/>= ls
        if (c.b.side != 0) return internal("couldn't parse as re >/>= ls");
        BRGH_TRY(set(&o->have_re, &o->re, c.a.column, "re"));
        BRGH_TRY(set(&o->have_ls, &o->ls, c.b.column, "ls"));
        o->re_minus_one = is_gt;
    } else if (c.a.side == 0 && (is_gteq || is_gt)) {   // le >/>= rs
        if (c.b.side != 1) return internal("couldn't parse as le >/>= rs");
        BRGH_TRY(set(&o->have_rs, &o->rs, c.b.column, "rs"));
        BRGH_TRY(set(&o->have_le, &o->le, c.a.column, "le"));
        o->le_minus_one = is_gt;
    } else {
        return internal("couldn't parse left side as neither 'rs </<=' nor 'ls </<=' nor 're >/>=' nor 'le >/>='");
    }
    return kOk;
}
// try_parse (intervals.rs:184-218)
static Err parse_join_filter(const std::string& text, const std::string& la, const std::string& ra, ColIntervalsOut* out) {
    FilterLexer lx(text);
    FilterCmp c1, c2;
    BRGH_TRY(parse_filter_cmp(lx, la, ra, &c1));
    if (!lx.eat_word("AND")) {
        lx.ws();
        return internal(lx.i >= lx.s.size() ? "filter expression is not an AND" : "filter expression is not an AND");
    }
    BRGH_TRY(parse_filter_cmp(lx, la, ra, &c2));
    lx.ws();
    if (lx.i < lx.s.size()) return internal("left side of AND is not a BinaryExpr");  // (a third conjunct nests the AND)
    BRGH_TRY(apply_filter_cmp(c1, out));
    BRGH_TRY(apply_filter_cmp(c2, out));
    if (!out->have_ls) return internal("left start bound not found");
    if (!out->have_le) return internal("left end bound not found");
    if (!out->have_rs) return internal("right start bound not found");
    if (!out->have_re) return internal("right end bound not found");
    return kOk;
}

}  // namespace

// ============================================================================================================================
// C entry points
// ============================================================================================================================
extern "C" {

BRG_API int brgh_create_bio_session(int device_id, brgh_session** out) {
    if (!out) return finish(internal("brgh_create_bio_session: out is NULL"));
    *out = new brgh_session();
    (*out)->device = device_id;
    return 0;
}
BRG_API void brgh_session_destroy(brgh_session* s) {
    if (!s) return;
    s->tables.clear();
    if (s->ctx) brg_ctx_destroy(s->ctx);
    delete s;
}
BRG_API int brgh_session_set(brgh_session* s, const char* key, const char* value) {
    if (!s || !key || !value) return finish(internal("brgh_session_set: NULL argument"));
    const std::string k = key, v = value;
    if (k == "bio.interval_join_algorithm") {
        const std::string l = lower(v);  // Algorithm::from_str (session_context.rs:74-90)
        for (int i = 0; i < 7; i++)
            if (l == lower(kAlgNames[i])) { s->algorithm = (Algorithm)i; return 0; }
        return finish(Err{BRGH_ERR_EXTERNAL, "Can't parse '" + v + "' as Algorithm"});
    }
    if (k == "bio.prefer_interval_join") return finish(parse_bool(v, &s->prefer_interval_join));
    if (k == "bio.interval_join_low_memory") return finish(parse_bool(v, &s->low_memory));
    return finish(Err{BRGH_ERR_EXTERNAL, "Config value \"" + k.substr(k.find('.') == std::string::npos ? 0 : k.find('.') + 1) +
                                          "\" not found on BioConfig"});
}
BRG_API int brgh_session_get(const brgh_session* s, const char* key, char* buf, uint64_t cap) {
    if (!s || !key || !buf || !cap) return finish(internal("brgh_session_get: NULL argument"));
    const std::string k = key;
    std::string v;
    if (k == "bio.interval_join_algorithm") v = kAlgNames[s->algorithm];
    else if (k == "bio.prefer_interval_join") v = s->prefer_interval_join ? "true" : "false";
    else if (k == "bio.interval_join_low_memory") v = s->low_memory ? "true" : "false";
    else return finish(Err{BRGH_ERR_EXTERNAL, "Config value \"" + k + "\" not found on BioConfig"});
    if (v.size() + 1 > cap) return finish(internal("brgh_session_get: buffer too small"));
    memcpy(buf, v.c_str(), v.size() + 1);
    return 0;
}
BRG_API int brgh_session_set_coalesce_rows(brgh_session* s, uint64_t rows) {
    if (!s || rows == 0) return finish(internal("brgh_session_set_coalesce_rows: bad argument"));
    s->coalesce_rows = rows;
    return 0;
}

BRG_API int brgh_register_batch(brgh_session* s, const char* table, ArrowArray* batch, ArrowSchema* schema) {
    if (!s || !table) return finish(internal("brgh_register_batch: NULL argument"));
    Batch b;
    std::vector<Field> sch;
    Err e = import_batch(batch, schema, &b, &sch);
    if (e.code) return finish(e);
    Table& t = s->tables[table];
    if (t.batches.empty() && t.schema.empty()) t.schema = sch;
    else {
        bool same = t.schema.size() == sch.size();
        for (size_t i = 0; same && i < sch.size(); i++) same = t.schema[i].name == sch[i].name && t.schema[i].format == sch[i].format;
        if (!same) return finish(plan(std::string("batch schema differs from the schema of table '") + table + "'"));
    }
    t.batches.push_back(std::move(b));
    return 0;
}
BRG_API int brgh_deregister_table(brgh_session* s, const char* table) {
    if (!s || !table) return finish(internal("brgh_deregister_table: NULL argument"));
    s->tables.erase(table);
    return 0;
}

static int run_to_result(Err e, brgh_result* r, brgh_result** out) {
    if (e.code) { delete r; return finish(e); }
    *out = r;
    return 0;
}
BRG_API int brgh_count_overlaps(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                                brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_count_overlaps: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_count(s, left, right, args, n_args, false, r), r, out);
}
BRG_API int brgh_coverage(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                          brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_coverage: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_count(s, left, right, args, n_args, true, r), r, out);
}
BRG_API int brgh_merge(brgh_session* s, const char* table, int64_t min_dist, const char* const* args, int n_args, brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_merge: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_merge(s, table, min_dist, args, n_args, false, r), r, out);
}
BRG_API int brgh_subtract(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                          brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_subtract: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_subtract(s, left, right, args, n_args, r), r, out);
}
BRG_API int brgh_complement(brgh_session* s, const char* table, const char* const* args, int n_args, brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_complement: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_complement(s, table, args, n_args, r), r, out);
}
BRG_API int brgh_cluster(brgh_session* s, const char* table, int64_t min_dist, const char* const* args, int n_args, brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_cluster: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_merge(s, table, min_dist, args, n_args, true, r), r, out);
}
BRG_API int brgh_nearest(brgh_session* s, const char* left, const char* right, int64_t k, int include_overlaps,
                         int compute_distance, const char* const* args, int n_args, brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_nearest: NULL argument"));
    auto* r = new brgh_result();
    return run_to_result(run_nearest(s, left, right, k, include_overlaps != 0, compute_distance != 0, args, n_args, r), r, out);
}
BRG_API int brgh_overlap(brgh_session* s, const char* left, const char* right, const char* const* args, int n_args,
                         brgh_result** out) {
    if (!s || !out) return finish(internal("brgh_overlap: NULL argument"));
    auto* r = new brgh_result();
    UdtfArgs ua;
    Err e = parse_overlap_args(args, n_args, "overlap", &ua);
    const Table *lt = nullptr, *rt = nullptr;
    if (!e.code) e = find_table(s, left, &lt);
    if (!e.code) e = find_table(s, right, &rt);
    if (!e.code) {
        // overlap.rs:152-226: FROM right AS b, left AS a WHERE a.contig = b.contig AND a.end >(=) b.start AND a.start <(=) b.end
        // => the RIGHT table is the SQL-left (build) input, the LEFT table is probed and defines the row order
        const char* onb[1] = {ua.right.contig.c_str()};
        const char* ona[1] = {ua.left.contig.c_str()};
        const bool join_mode = ua.output_mode == BRGH_OVERLAP_JOIN;
        const int jt = ua.output_mode == BRGH_OVERLAP_LEFT_DISTINCT ? BRGH_JOIN_RIGHT_SEMI : BRGH_JOIN_INNER;
        // build = b (right table): b.start <= a.end [strict: <  -> a.end - 1], b.end >= a.start [strict: > -> b.end - 1]
        // output column order: left_* then right_* (select_left, select_right)
        brgh_result tmp;
        e = run_join(s, *rt, *lt, onb, ona, 1, ua.right.start, ua.right.end, ua.left.start, ua.left.end, ua.strict, ua.strict, jt,
                     join_mode ? "right_" : "", join_mode ? "left_" : "", join_mode, true, &tmp);
        if (!e.code) {
            for (auto& b : tmp.batches) {
                if (join_mode) {  // run_join emitted build (right_*) columns first: the UDTF's schema is left_* ++ right_*
                    const size_t nb = rt->schema.size();
                    std::rotate(b.cols.begin(), b.cols.begin() + (long)nb, b.cols.end());
                }
                r->batches.push_back(std::move(b));
            }
        }
    }
    return run_to_result(e, r, out);
}
BRG_API int brgh_interval_join(brgh_session* s, const char* left, const char* right, const char* const* on_left,
                               const char* const* on_right, int n_on, const char* left_start, const char* left_end,
                               const char* right_start, const char* right_end, int lt_strict, int gt_strict, int join_type,
                               brgh_result** out) {
    if (!s || !out || n_on < 0 || !left_start || !left_end || !right_start || !right_end)
        return finish(internal("brgh_interval_join: bad argument"));
    auto* r = new brgh_result();
    const Table *lt = nullptr, *rt = nullptr;
    Err e = find_table(s, left, &lt);
    if (!e.code) e = find_table(s, right, &rt);
    if (!e.code)
        e = run_join(s, *lt, *rt, on_left, on_right, n_on, left_start, left_end, right_start, right_end, lt_strict != 0, gt_strict != 0,
                     join_type, "", "", true, true, r);
    return run_to_result(e, r, out);
}
BRG_API int brgh_parse_join_filter(const char* filter, const char* left_alias, const char* right_alias, brgh_col_intervals* out) {
    if (!filter || !left_alias || !right_alias || !out) return finish(internal("brgh_parse_join_filter: NULL argument"));
    ColIntervalsOut o;
    Err e = parse_join_filter(filter, left_alias, right_alias, &o);
    if (e.code) return finish(e);
    auto put = [&](char* dst, const std::string& v) -> bool {
        if (v.size() + 1 > sizeof(out->left_start)) return false;
        memcpy(dst, v.c_str(), v.size() + 1);
        return true;
    };
    if (!put(out->left_start, o.ls) || !put(out->left_end, o.le) || !put(out->right_start, o.rs) || !put(out->right_end, o.re))
        return finish(internal("brgh_parse_join_filter: column name too long"));
    out->left_end_minus_one = o.le_minus_one ? 1 : 0;
    out->right_end_minus_one = o.re_minus_one ? 1 : 0;
    return 0;
}
BRG_API int brgh_interval_join_filter(brgh_session* s, const char* left, const char* right, const char* const* on_left,
                                      const char* const* on_right, int n_on, const char* filter, const char* left_alias,
                                      const char* right_alias, int join_type, brgh_result** out) {
    if (!s || !left || !right || !filter || !out) return finish(internal("brgh_interval_join_filter: NULL argument"));
    brgh_col_intervals ci;
    const int rc = brgh_parse_join_filter(filter, left_alias ? left_alias : left, right_alias ? right_alias : right, &ci);
    if (rc) return rc;
    // `l.end > r.start` compares l.end - 1 (gt_strict), `l.start < r.end` compares r.end - 1 (lt_strict): intervals.rs:85-115
    return brgh_interval_join(s, left, right, on_left, on_right, n_on, ci.left_start, ci.left_end, ci.right_start, ci.right_end,
                              ci.right_end_minus_one, ci.left_end_minus_one, join_type, out);
}
BRG_API int brgh_explain_interval_join(const brgh_session* s, int join_type, char* buf, uint64_t cap) {
    if (!s || !buf || !cap) return finish(internal("brgh_explain_interval_join: NULL argument"));
    const char* jt = join_type == BRGH_JOIN_RIGHT_SEMI ? "RightSemi" : (join_type == BRGH_JOIN_RIGHT_ANTI ? "RightAnti" : "Inner");
    const std::string v = std::string("IntervalJoinExec: mode=CollectLeft, join_type=") + jt + ", alg=" + kAlgNames[s->algorithm];
    if (v.size() + 1 > cap) return finish(internal("brgh_explain_interval_join: buffer too small"));
    memcpy(buf, v.c_str(), v.size() + 1);
    return 0;
}

BRG_API uint64_t brgh_result_num_batches(const brgh_result* r) { return r ? r->batches.size() : 0; }
BRG_API uint64_t brgh_result_num_rows(const brgh_result* r) {
    uint64_t n = 0;
    if (r) for (auto& b : r->batches) n += (uint64_t)b.rows;
    return n;
}
BRG_API int brgh_result_export(brgh_result* r, uint64_t i, ArrowArray* out, ArrowSchema* out_schema) {
    if (!r || i >= r->batches.size() || !out || !out_schema) return finish(internal("brgh_result_export: bad argument"));
    export_batch(r->batches[i], out);
    export_schema(r->batches[i], out_schema);
    return 0;
}
BRG_API void brgh_result_release(brgh_result* r) { delete r; }

BRG_API int brgh_resolve_positions(ArrowArray* batch, ArrowSchema* schema, const char* name, const char* label, int32_t* out,
                                   uint64_t cap) {
    Batch b;
    Err e = import_batch(batch, schema, &b, nullptr);
    if (e.code) return finish(e);
    std::vector<int32_t> v;
    e = resolve_positions(b, name ? name : "", label ? label : "start", &v);
    if (e.code) return finish(e);
    if (v.size() > cap) return finish(internal("brgh_resolve_positions: output buffer too small"));
    if (!v.empty()) memcpy(out, v.data(), v.size() * 4);
    return 0;
}
BRG_API int brgh_take(ArrowArray* batch, ArrowSchema* schema, const uint32_t* idx, const uint8_t* idx_valid_bits, uint64_t n,
                      const char* name_prefix, ArrowArray* out, ArrowSchema* out_schema) {
    Table t;
    Batch b;
    Err e = import_batch(batch, schema, &b, &t.schema);
    if (e.code) return finish(e);
    for (uint64_t i = 0; i < n; i++) {
        const bool null = idx_valid_bits && !((idx_valid_bits[i >> 3] >> (i & 7)) & 1);
        if (!null && (int64_t)idx[i] >= b.rows)
            return finish(Err{BRGH_ERR_ARROW, "Arrow error: Compute error: Array index out of bounds, cannot get item at index " +
                                                  std::to_string(idx[i]) + " from " + std::to_string(b.rows) + " entries"});
    }
    t.batches.push_back(std::move(b));
    Batch ob;
    ob.rows = (int64_t)n;
    e = take_table(t, idx, idx_valid_bits, (int64_t)n, name_prefix ? name_prefix : "", -1, &ob.cols);
    if (e.code) return finish(e);
    export_batch(ob, out);
    export_schema(ob, out_schema);
    return 0;
}

}  // extern "C"
