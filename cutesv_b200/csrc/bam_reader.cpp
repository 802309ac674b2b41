// bam_reader.cpp -- native BAM -> columnar packet decoder (host data loader).
//
// Replaces the per-read Python loop over pysam objects (cuteSV:709-733) for plain BAM input: BGZF
// blocks are inflated by a small thread pool, records are parsed straight into the int32 columns /
// BAM-native u32 CIGAR array / reduced SA-segment table that csv_extract consumes (see
// include/cutesv_b200.h: csv_read_cols, csv_sa_cols).  No htslib: the BAM and BGZF layouts are
// implemented from the SAM/BAM specification.  Built as libcutesv_bam.so (g++ -lz -pthread).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

struct Block { std::vector<uint8_t> comp; std::vector<uint8_t> raw; uint32_t isize; bool ok; };

bool inflate_block(Block& b) {
    b.raw.resize(b.isize);
    if (b.isize == 0) return true;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = b.comp.data();
    zs.avail_in = (uInt)b.comp.size();
    zs.next_out = b.raw.data();
    zs.avail_out = (uInt)b.raw.size();
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.total_out == b.isize;
}

// A batch of BGZF blocks being inflated by the pool.
struct Batch {
    std::vector<Block> blocks;
    size_t next = 0, done = 0;   // guarded by Pool::m
};

// Persistent inflate workers (threads that live as long as the reader: short-lived threads do not get
// spread over the cores quickly enough for 64 KB blocks).
struct Pool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    Batch* cur = nullptr;
    bool stop = false;
    void start(int n) {
        for (int i = 0; i < n; i++) th.emplace_back([this]() { run(); });
    }
    void run() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_work.wait(lk, [this]() { return stop || (cur && cur->next < cur->blocks.size()); });
            if (stop) return;
            Batch* b = cur;
            const size_t i = b->next++;
            lk.unlock();
            b->blocks[i].ok = inflate_block(b->blocks[i]);
            lk.lock();
            if (++b->done == b->blocks.size()) cv_done.notify_all();
        }
    }
    void submit(Batch* b) {
        { std::lock_guard<std::mutex> lk(m); cur = b; }
        cv_work.notify_all();
    }
    void wait(Batch* b) {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [b]() { return b->done == b->blocks.size(); });
        if (cur == b) cur = nullptr;
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_work.notify_all();
        for (auto& t : th) t.join();
    }
};

struct Reader {
    FILE* f = nullptr;
    bool eof = false;            // no more blocks in the file
    std::vector<uint8_t> buf;    // decompressed stream not yet consumed
    size_t pos = 0;
    int n_threads = 4;
    Pool pool;
    Batch batch[2];              // [inflight] is being inflated while the parser consumes the other's payload
    int inflight = -1;
    std::string io_err;          // error met while reading ahead (reported when that batch is consumed)
    // header
    std::vector<std::string> ref_name;
    std::vector<int64_t> ref_len;
    std::vector<int32_t> chrom_id;   // header index -> contig id handed to the kernels
    std::unordered_map<std::string, int32_t> ref_index;
    // read names -> provisional ids (first-seen order)
    std::unordered_map<std::string, int32_t> name_id;
    std::vector<std::string> names;
    // packet storage (valid until the next bamr_next)
    std::vector<int32_t> chrom, ref_start, ref_end, flag, mapq, query_len, read_id;
    std::vector<int64_t> cigar_off, sa_off, seq_off;
    std::vector<uint32_t> cigar;
    std::vector<int32_t> sa_chrom, sa_pos0, sa_strand, sa_mapq, sa_first, sa_last, sa_span;
    std::vector<uint8_t> seq4;
    bool keep_seq = true;
};

// read up to `max_blocks` raw BGZF blocks from the file into `b`; false + r.io_err on a malformed file
bool read_blocks(Reader& r, Batch& b, int max_blocks) {
    b.blocks.clear();
    b.next = b.done = 0;
    for (int i = 0; i < max_blocks && !r.eof; i++) {
        uint8_t h[18];
        size_t got = fread(h, 1, 18, r.f);
        if (got == 0) { r.eof = true; break; }
        if (got != 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { r.io_err = got == 18 ? "not a BGZF block" : "truncated BGZF header"; return false; }
        const int xlen = h[10] | (h[11] << 8);
        int bsize = -1;
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), h + 12, std::min(6, xlen));
        if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, r.f) != (size_t)(xlen - 6)) { r.io_err = "truncated BGZF header"; return false; }
        for (int o = 0; o + 4 <= xlen;) {
            const int slen = extra[o + 2] | (extra[o + 3] << 8);
            if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = extra[o + 4] | (extra[o + 5] << 8);
            o += 4 + slen;
        }
        if (bsize < 0) { r.io_err = "BGZF block without BC field"; return false; }
        const int clen = bsize + 1 - 12 - xlen - 8;  // deflate payload length
        Block blk;
        blk.comp.resize(clen > 0 ? clen : 0);
        uint8_t tail[8];
        if ((clen > 0 && fread(blk.comp.data(), 1, clen, r.f) != (size_t)clen) || fread(tail, 1, 8, r.f) != 8) { r.io_err = "truncated BGZF block"; return false; }
        blk.isize = tail[4] | (tail[5] << 8) | (tail[6] << 16) | ((uint32_t)tail[7] << 24);
        blk.ok = true;
        b.blocks.push_back(std::move(blk));
    }
    return true;
}

// start inflating the next batch of the file (if any) in the background
void prefetch(Reader& r) {
    if (r.inflight >= 0 || r.eof || !r.io_err.empty()) return;
    const int slot = 0;  // batch[0] is always the in-flight one; its payload is moved into buf on arrival
    read_blocks(r, r.batch[slot], 256);
    if (r.batch[slot].blocks.empty()) return;
    r.inflight = slot;
    r.pool.submit(&r.batch[slot]);
}

// append the payload of the next batch to r.buf; false at end of file or on error (g_err set)
bool refill(Reader& r) {
    if (r.pos > 0) {  // drop the consumed prefix
        r.buf.erase(r.buf.begin(), r.buf.begin() + (long)r.pos);
        r.pos = 0;
    }
    if (r.inflight < 0) prefetch(r);
    if (r.inflight < 0) {
        if (!r.io_err.empty()) g_err = r.io_err;
        return false;
    }
    Batch& b = r.batch[r.inflight];
    r.pool.wait(&b);
    size_t total = 0;
    for (auto& blk : b.blocks) {
        if (!blk.ok) { g_err = "BGZF inflate failed"; return false; }
        total += blk.raw.size();
    }
    size_t o = r.buf.size();
    r.buf.resize(o + total);
    for (auto& blk : b.blocks) {
        if (!blk.raw.empty()) memcpy(r.buf.data() + o, blk.raw.data(), blk.raw.size());
        o += blk.raw.size();
    }
    r.inflight = -1;
    prefetch(r);   // the next batch inflates while the caller parses this one
    return true;
}

// make sure `need` bytes are available at r.pos; false at clean EOF
bool ensure(Reader& r, size_t need) {
    while (r.buf.size() - r.pos < need) {
        if (!refill(r)) return false;
    }
    return true;
}

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

// (first S length, last S length, reference span) of a CIGAR string: acquire_clip_pos, cuteSV:466-481
void clip_pos(const char* s, size_t n, int32_t* first, int32_t* last, int32_t* span) {
    *first = 0; *last = 0; *span = 0;
    int64_t num = 0;
    bool first_op = true;
    int32_t last_len = 0;
    char last_op = 0;
    for (size_t i = 0; i < n; i++) {
        const char c = s[i];
        if (c >= '0' && c <= '9') { num = num * 10 + (c - '0'); continue; }
        if (first_op) { if (c == 'S') *first = (int32_t)num; first_op = false; }
        if (c == 'M' || c == 'D' || c == '=' || c == 'X') *span += (int32_t)num;
        last_len = (int32_t)num; last_op = c;
        num = 0;
    }
    if (last_op == 'S') *last = last_len;
}

size_t aux_skip(const uint8_t* p, const uint8_t* end, char type) {
    switch (type) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'Z': case 'H': { size_t n = 0; while (p + n < end && p[n]) n++; return n + 1; }
        case 'B': {
            if (p + 5 > end) return (size_t)(end - p);
            const char st = (char)p[0];
            const uint32_t cnt = rd_u32(p + 1);
            const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            return 5 + es * cnt;
        }
        default: return (size_t)(end - p);
    }
}

}  // namespace

extern "C" {

typedef struct bamr_packet {
    int64_t n;
    int32_t *chrom, *ref_start, *ref_end, *flag, *mapq, *query_len, *read_id;
    int64_t *cigar_off, *sa_off;
    int64_t n_cigar;
    uint32_t* cigar;
    int64_t n_sa;
    int32_t *sa_chrom, *sa_pos0, *sa_strand, *sa_mapq, *sa_first, *sa_last, *sa_span;
    int64_t* seq_off;
    uint8_t* seq4;
} bamr_packet;

const char* bamr_error(void) { return g_err.c_str(); }

int bamr_open(const char* path, int n_threads, void** out) {
    Reader* r = new Reader();
    r->f = fopen(path, "rb");
    if (!r->f) { g_err = std::string("cannot open ") + path; delete r; return -1; }
    r->n_threads = n_threads > 0 ? n_threads : 4;
    r->pool.start(r->n_threads);
    g_err.clear();
    if (!ensure(*r, 12) || memcmp(r->buf.data() + r->pos, "BAM\1", 4) != 0) {
        if (g_err.empty()) g_err = "not a BAM file";
        fclose(r->f); delete r; return -1;
    }
    const int32_t l_text = rd_i32(r->buf.data() + r->pos + 4);
    if (!ensure(*r, 12 + (size_t)l_text)) { g_err = "truncated BAM header"; fclose(r->f); delete r; return -1; }
    r->pos += 8 + (size_t)l_text;
    const int32_t n_ref = rd_i32(r->buf.data() + r->pos);
    r->pos += 4;
    for (int32_t i = 0; i < n_ref; i++) {
        if (!ensure(*r, 4)) { g_err = "truncated BAM header"; fclose(r->f); delete r; return -1; }
        const int32_t l_name = rd_i32(r->buf.data() + r->pos);
        if (!ensure(*r, 4 + (size_t)l_name + 4)) { g_err = "truncated BAM header"; fclose(r->f); delete r; return -1; }
        std::string nm((const char*)r->buf.data() + r->pos + 4, (size_t)l_name - 1);
        const int32_t l_ref = rd_i32(r->buf.data() + r->pos + 4 + l_name);
        r->pos += 8 + (size_t)l_name;
        r->ref_index[nm] = i;
        r->ref_name.push_back(nm);
        r->ref_len.push_back(l_ref);
        r->chrom_id.push_back(i);
    }
    *out = r;
    return 0;
}

void bamr_close(void* h) {
    Reader* r = (Reader*)h;
    if (!r) return;
    if (r->f) fclose(r->f);
    delete r;
}

int32_t bamr_n_ref(void* h) { return (int32_t)((Reader*)h)->ref_name.size(); }
const char* bamr_ref_name(void* h, int32_t i) { return ((Reader*)h)->ref_name[i].c_str(); }
int64_t bamr_ref_len(void* h, int32_t i) { return ((Reader*)h)->ref_len[i]; }
void bamr_set_chrom_ids(void* h, const int32_t* ids) { Reader* r = (Reader*)h; for (size_t i = 0; i < r->chrom_id.size(); i++) r->chrom_id[i] = ids[i]; }
void bamr_keep_seq(void* h, int keep) { ((Reader*)h)->keep_seq = keep != 0; }

// Next packet of up to max_records MAPPED records (file order).  Returns the record count (0 at EOF,
// -1 on error); pointers stay valid until the next call.
int64_t bamr_next(void* h, int64_t max_records, bamr_packet* out) {
    Reader& r = *(Reader*)h;
    r.chrom.clear(); r.ref_start.clear(); r.ref_end.clear(); r.flag.clear(); r.mapq.clear(); r.query_len.clear(); r.read_id.clear();
    r.cigar_off.assign(1, 0); r.sa_off.assign(1, 0); r.seq_off.assign(1, 0);
    r.cigar.clear(); r.seq4.clear();
    r.sa_chrom.clear(); r.sa_pos0.clear(); r.sa_strand.clear(); r.sa_mapq.clear(); r.sa_first.clear(); r.sa_last.clear(); r.sa_span.clear();
    int64_t n = 0;
    while (n < max_records) {
        if (!ensure(r, 4)) break;
        const int32_t block_size = rd_i32(r.buf.data() + r.pos);
        if (block_size < 32) { g_err = "corrupt BAM record"; return -1; }
        if (!ensure(r, 4 + (size_t)block_size)) { g_err = "truncated BAM record"; return -1; }
        const uint8_t* p = r.buf.data() + r.pos + 4;
        const uint8_t* end = p + block_size;
        r.pos += 4 + (size_t)block_size;
        const int32_t ref_id = rd_i32(p), pos = rd_i32(p + 4);
        const int l_read_name = p[8], mq = p[9];
        const int n_cigar = rd_u16(p + 12), flg = rd_u16(p + 14);
        const int32_t l_seq = rd_i32(p + 16);
        if (ref_id < 0 || ref_id >= (int32_t)r.ref_name.size()) continue;  // unmapped / unplaced: never returned by fetch(chr, ...)
        const char* qname = (const char*)p + 32;
        const uint8_t* cg = p + 32 + l_read_name;
        const uint8_t* sq = cg + 4 * (size_t)n_cigar;
        const uint8_t* aux = sq + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
        // tags: SA:Z and the long-CIGAR CG:B,I
        const char* sa = nullptr;
        const uint8_t* cg_tag = nullptr;
        uint32_t cg_cnt = 0;
        for (const uint8_t* a = aux; a + 3 <= end;) {
            const char t0 = (char)a[0], t1 = (char)a[1], ty = (char)a[2];
            const uint8_t* v = a + 3;
            if (t0 == 'S' && t1 == 'A' && ty == 'Z') sa = (const char*)v;
            if (t0 == 'C' && t1 == 'G' && ty == 'B' && v + 5 <= end && (v[0] == 'I' || v[0] == 'i')) { cg_cnt = rd_u32(v + 1); cg_tag = v + 5; }
            a = v + aux_skip(v, end, ty);
        }
        const uint8_t* cig_src = cg;
        uint32_t cig_n = (uint32_t)n_cigar;
        if (cg_tag && n_cigar == 2 && (rd_u32(cg) & 15) == 4 && (int32_t)(rd_u32(cg) >> 4) == l_seq && (rd_u32(cg + 4) & 15) == 3) {
            cig_src = cg_tag; cig_n = cg_cnt;  // real CIGAR of a >65535-op alignment lives in the CG tag
        }
        int32_t span = 0;
        for (uint32_t k = 0; k < cig_n; k++) {
            const uint32_t c = rd_u32(cig_src + 4 * (size_t)k);
            r.cigar.push_back(c);
            const int op = c & 15;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += (int32_t)(c >> 4);
        }
        r.cigar_off.push_back((int64_t)r.cigar.size());
        std::string name(qname, (size_t)(l_read_name > 0 ? l_read_name - 1 : 0));
        auto it = r.name_id.find(name);
        int32_t id;
        if (it == r.name_id.end()) { id = (int32_t)r.names.size(); r.name_id.emplace(name, id); r.names.push_back(name); }
        else id = it->second;
        r.chrom.push_back(r.chrom_id[ref_id]); r.ref_start.push_back(pos); r.ref_end.push_back(pos + span); r.flag.push_back(flg);
        r.mapq.push_back(mq); r.query_len.push_back(l_seq); r.read_id.push_back(id);
        if (sa) {  // "rname,pos,strand,CIGAR,mapQ,NM;" ... (cuteSV:489-509)
            const char* s = sa;
            while (*s) {
                const char* e = s;
                while (*e && *e != ';') e++;
                const char* f[6]; int nf = 0; f[nf++] = s;
                for (const char* q = s; q < e && nf < 6; q++) if (*q == ',') f[nf++] = q + 1;
                if (nf >= 5) {
                    std::string rn(f[0], (size_t)(f[1] - f[0] - 1));
                    auto ri = r.ref_index.find(rn);
                    int32_t fc, lc, sp;
                    clip_pos(f[3], (size_t)(f[4] - f[3] - 1), &fc, &lc, &sp);
                    r.sa_chrom.push_back(ri == r.ref_index.end() ? -1 : r.chrom_id[ri->second]);
                    r.sa_pos0.push_back(atoi(f[1]) - 1);
                    r.sa_strand.push_back(*f[2] == '+' ? 0 : 1);
                    r.sa_mapq.push_back(atoi(f[4]));
                    r.sa_first.push_back(fc); r.sa_last.push_back(lc); r.sa_span.push_back(sp);
                }
                s = *e ? e + 1 : e;
            }
        }
        r.sa_off.push_back((int64_t)r.sa_chrom.size());
        if (r.keep_seq) r.seq4.insert(r.seq4.end(), sq, sq + (size_t)(l_seq + 1) / 2);
        r.seq_off.push_back((int64_t)r.seq4.size());
        n++;
    }
    out->n = n;
    out->chrom = r.chrom.data(); out->ref_start = r.ref_start.data(); out->ref_end = r.ref_end.data(); out->flag = r.flag.data();
    out->mapq = r.mapq.data(); out->query_len = r.query_len.data(); out->read_id = r.read_id.data();
    out->cigar_off = r.cigar_off.data(); out->sa_off = r.sa_off.data();
    out->n_cigar = (int64_t)r.cigar.size(); out->cigar = r.cigar.data();
    out->n_sa = (int64_t)r.sa_chrom.size();
    out->sa_chrom = r.sa_chrom.data(); out->sa_pos0 = r.sa_pos0.data(); out->sa_strand = r.sa_strand.data(); out->sa_mapq = r.sa_mapq.data();
    out->sa_first = r.sa_first.data(); out->sa_last = r.sa_last.data(); out->sa_span = r.sa_span.data();
    out->seq_off = r.seq_off.data(); out->seq4 = r.seq4.data();
    return n;
}

int64_t bamr_n_names(void* h) { return (int64_t)((Reader*)h)->names.size(); }
const char* bamr_name(void* h, int64_t id) { return ((Reader*)h)->names[(size_t)id].c_str(); }
// rank[provisional id] = rank of the name in byte-wise (== Python str for ASCII) order
void bamr_name_ranks(void* h, int32_t* rank) {
    Reader& r = *(Reader*)h;
    std::vector<int32_t> order(r.names.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int32_t)i;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return r.names[a] < r.names[b]; });
    for (size_t i = 0; i < order.size(); i++) rank[order[i]] = (int32_t)i;
}
// 4-bit packed -> ASCII
void bamr_decode_seq(const uint8_t* seq4, int64_t l_seq, char* out) {
    static const char tab[] = "=ACMGRSVTWYHKDBN";
    for (int64_t i = 0; i < l_seq; i++) out[i] = tab[(seq4[i >> 1] >> ((~i & 1) << 2)) & 15];
}

// Per-reference mapped-read counts from a .bai index (get_index_statistics, cuteSV:1015-1025):
// the pseudo-bin 37450 of every reference holds (n_mapped, n_unmapped).
int bamr_index_stats(const char* bai_path, int32_t n_ref, int64_t* mapped) {
    FILE* f = fopen(bai_path, "rb");
    if (!f) { g_err = std::string("cannot open ") + bai_path; return -1; }
    uint8_t h[8];
    if (fread(h, 1, 8, f) != 8 || memcmp(h, "BAI\1", 4) != 0) { fclose(f); g_err = "not a BAI index"; return -1; }
    const int32_t n = rd_i32(h + 4);
    for (int32_t i = 0; i < n_ref; i++) mapped[i] = 0;
    for (int32_t ref = 0; ref < n; ref++) {
        uint8_t b4[4];
        if (fread(b4, 1, 4, f) != 4) break;
        const int32_t n_bin = rd_i32(b4);
        for (int32_t b = 0; b < n_bin; b++) {
            uint8_t bh[8];
            if (fread(bh, 1, 8, f) != 8) { fclose(f); g_err = "truncated BAI"; return -1; }
            const uint32_t bin = rd_u32(bh);
            const int32_t n_chunk = rd_i32(bh + 4);
            std::vector<uint8_t> ch((size_t)n_chunk * 16);
            if (n_chunk && fread(ch.data(), 1, ch.size(), f) != ch.size()) { fclose(f); g_err = "truncated BAI"; return -1; }
            if (bin == 37450 && n_chunk >= 2 && ref < n_ref) { uint64_t m; memcpy(&m, ch.data() + 16, 8); mapped[ref] = (int64_t)m; }
        }
        if (fread(b4, 1, 4, f) != 4) break;
        const int32_t n_intv = rd_i32(b4);
        if (n_intv && fseek(f, (long)n_intv * 8, SEEK_CUR) != 0) break;
    }
    fclose(f);
    return 0;
}

}  // extern "C"
