// extract_api.inl -- C-ABI of the extraction stage (included by cutesv_b200.cu)

static int ex_upload(csv_ctx* c, DBuf& b, const void* src, size_t bytes) {
    CU(b.ensure(bytes ? bytes : 4));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return CSV_OK;
}
// capacity for `bytes`, keeping the first keep_bytes of the present contents (append mode)
static int ex_grow(csv_ctx* c, DBuf& b, size_t bytes, size_t keep_bytes) {
    if (bytes <= b.cap) return CSV_OK;
    if (keep_bytes == 0 || !b.p) { CU(b.ensure(bytes)); return CSV_OK; }
    DBuf nb;
    CU(nb.ensure(bytes + bytes / 2));   // geometric growth: a run appends many packets
    CU(cudaMemcpyAsync(nb.p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    b.release();
    b = nb;
    return CSV_OK;
}

// read ids of everything extracted so far: id -> rank[id] (the CLI numbers read names in first-seen order while it decodes
// and only knows their ranks in Python string order at the end, cuteSV:764-801)
__global__ void k_remap_ids(int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ rank, int64_t n_rank, uint32_t* status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = ids[i];
        if (v < 0 || v >= n_rank) { atomicOr(status, ST_NEG_FIELD); continue; }
        ids[i] = rank[v];
    }
}

static int extract_impl(csv_ctx* c, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar, const csv_sa_cols* sa,
                        int64_t counts[CSV_NTYPES], int64_t* n_read_rows, bool append) {
    if (!c || !reads) return set_err(CSV_E_INVALID, "null argument");
    if (c->n_contigs == 0) return set_err(CSV_E_STATE, "csv_set_contigs has not been called");
    const int64_t n = reads->n;
    if (n < 0 || n >= (1ll << 30)) return set_err(CSV_E_INVALID, "record count out of range");
    CU(cudaSetDevice(c->device));
    ExtractState& X = c->ex;
    const int64_t n_sa = sa ? sa->n : 0;
    for (int slot = 0; slot <= CSV_NTYPES; slot++) {  // extraction writes the device-resident inputs: drain pending uploads first
        int wrc = wait_upload(c, slot);
        if (wrc) return wrc;
    }
    if (!append || !X.appending) {   // a fresh accumulation
        for (int t = 0; t < CSV_NTYPES; t++) c->sig[t].n = 0;
        c->n_reads = 0;
        X.n_pieces = 0; X.n_records = 0; X.n_skipped = 0;
    }
    X.appending = append;
    uint32_t base[8];   // counters at the start of this packet: signatures per type, pieces, reads rows, (status)
    for (int t = 0; t < CSV_NTYPES; t++) base[t] = (uint32_t)c->sig[t].n;
    base[5] = X.n_pieces; base[6] = (uint32_t)c->n_reads; base[7] = 0;
    if ((int64_t)X.n_records + n >= (1ll << 31)) return set_err(CSV_E_INVALID, "more than 2^31 alignment records in one accumulation");
    stage_begin(c, CSV_ST_H2D);
    int rc;
    const void* rsrc[7] = {reads->chrom, reads->ref_start, reads->ref_end, reads->flag, reads->mapq, reads->query_len, reads->read_id};
    for (int k = 0; k < 7; k++) { rc = ex_upload(c, X.r[k], rsrc[k], (size_t)n * 4); if (rc) return rc; }
    rc = ex_upload(c, X.cigar_off, reads->cigar_off, (size_t)(n + 1) * 8); if (rc) return rc;
    rc = ex_upload(c, X.sa_off, reads->sa_off, (size_t)(n + 1) * 8); if (rc) return rc;
    CU(X.cigar.ensure((size_t)n_cigar * 4 + 2 * EX_SLOT_BYTES));   // the bulk copies of k_extract read whole tiles
    rc = ex_upload(c, X.cigar, cigar, (size_t)n_cigar * 4); if (rc) return rc;
    const void* ssrc[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (n_sa) { ssrc[0] = sa->chrom; ssrc[1] = sa->pos0; ssrc[2] = sa->strand; ssrc[3] = sa->mapq; ssrc[4] = sa->first_clip; ssrc[5] = sa->last_clip; ssrc[6] = sa->ref_span; }
    for (int k = 0; k < 7; k++) { rc = ex_upload(c, X.s[k], ssrc[k], (size_t)n_sa * 4); if (rc) return rc; }
    stage_end(c, CSV_ST_H2D);
    CU(X.counters.ensure(16 * 4));
    if (!X.h_counters) CU(cudaMallocHost((void**)&X.h_counters, 32 * 4));
    // first guess of the room this packet needs; the kernel keeps counting past the capacities, so one rerun suffices
    uint32_t cap[CSV_NTYPES], cap_pieces, cap_rows = base[6] + (uint32_t)n + 16;
    const int64_t lim = (1ll << 30) - 1;
    cap[CSV_DEL] = (uint32_t)std::min<int64_t>((int64_t)base[CSV_DEL] + 4 * n + 1024, lim);
    cap[CSV_INS] = (uint32_t)std::min<int64_t>((int64_t)base[CSV_INS] + 4 * n + 1024, lim);
    for (int t = CSV_INV; t < CSV_NTYPES; t++) cap[t] = (uint32_t)std::min<int64_t>((int64_t)base[t] + 2 * n_sa + n / 4 + 1024, lim);
    cap_pieces = (uint32_t)std::min<int64_t>((int64_t)base[5] + 8 * n + 2048, 2 * lim);
    // later packets of a run: what the densest packet so far yielded per record, with head room (a rerun costs a second pass
    // over the packet and a reallocation of every output column)
    for (int t = 0; t < CSV_NTYPES; t++)
        cap[t] = (uint32_t)std::min<int64_t>(std::max<int64_t>(cap[t], (int64_t)base[t] + (int64_t)(1.25 * X.per_record[t] * (double)n) + 1024), lim);
    cap_pieces = (uint32_t)std::min<int64_t>(std::max<int64_t>(cap_pieces, (int64_t)base[5] + (int64_t)(1.25 * X.per_record[5] * (double)n) + 2048), 2 * lim);
    for (int attempt = 0; attempt < 3; attempt++) {
        ExtractOut O;
        memset(&O, 0, sizeof(O));
        for (int t = 0; t < CSV_NTYPES; t++) {
            SigBuf& sb = c->sig[t];
            const size_t keep = (size_t)base[t] * 4, want = (size_t)cap[t] * 4;
            if ((rc = ex_grow(c, sb.chrom, want, keep)) || (rc = ex_grow(c, sb.a, want, keep)) || (rc = ex_grow(c, sb.b, want, keep)) ||
                (rc = ex_grow(c, sb.rid, want, keep)) || (rc = ex_grow(c, sb.c, want, keep))) return rc;
            O.col[t][0] = sb.chrom.as<int32_t>(); O.col[t][1] = sb.a.as<int32_t>(); O.col[t][2] = sb.b.as<int32_t>();
            O.col[t][3] = sb.rid.as<int32_t>(); O.col[t][4] = sb.c.as<int32_t>();
            O.cap_sig[t] = cap[t];
        }
        if ((rc = ex_grow(c, X.piece_off, (size_t)cap[CSV_INS] * 4, (size_t)base[CSV_INS] * 4)) ||
            (rc = ex_grow(c, X.piece_cnt, (size_t)cap[CSV_INS] * 4, (size_t)base[CSV_INS] * 4)) ||
            (rc = ex_grow(c, X.pieces, (size_t)cap_pieces * sizeof(InsPiece), (size_t)base[5] * sizeof(InsPiece)))) return rc;
        const size_t keep_r = (size_t)base[6] * 4;
        if ((rc = ex_grow(c, c->r_chrom, (size_t)cap_rows * 4, keep_r)) || (rc = ex_grow(c, c->r_start, (size_t)cap_rows * 4, keep_r)) ||
            (rc = ex_grow(c, c->r_end, (size_t)cap_rows * 4, keep_r)) || (rc = ex_grow(c, c->r_id, (size_t)cap_rows * 4, keep_r)) ||
            (rc = ex_grow(c, c->r_prim, (size_t)cap_rows, (size_t)base[6]))) return rc;
        uint32_t* dc = X.counters.as<uint32_t>();
        O.n_sig = dc; O.n_pieces = dc + 5; O.n_rows = dc + 6; O.status = dc + 7; O.n_skipped = dc + 8;
        O.ins_piece_off = X.piece_off.as<int32_t>(); O.ins_piece_cnt = X.piece_cnt.as<int32_t>(); O.pieces = X.pieces.as<InsPiece>();
        O.cap_pieces = cap_pieces;
        O.rr_chrom = c->r_chrom.as<int32_t>(); O.rr_start = c->r_start.as<int32_t>(); O.rr_end = c->r_end.as<int32_t>();
        O.rr_id = c->r_id.as<int32_t>(); O.rr_prim = c->r_prim.as<uint8_t>(); O.cap_rows = cap_rows;
        // the append counters start at the totals so far (pinned staging: a rerun must not race the previous copy)
        uint32_t* hb = X.h_counters + 16;
        CU(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 16; k++) hb[k] = k < 8 ? base[k] : 0u;
        CU(cudaMemcpyAsync(dc, hb, 16 * 4, cudaMemcpyHostToDevice, c->stream));
        ReadView R;
        R.chrom = X.r[0].as<int32_t>(); R.ref_start = X.r[1].as<int32_t>(); R.ref_end = X.r[2].as<int32_t>(); R.flag = X.r[3].as<int32_t>();
        R.mapq = X.r[4].as<int32_t>(); R.query_len = X.r[5].as<int32_t>(); R.read_id = X.r[6].as<int32_t>();
        R.cigar_off = X.cigar_off.as<int64_t>(); R.sa_off = X.sa_off.as<int64_t>(); R.n = n;
        SaView S{X.s[0].as<int32_t>(), X.s[1].as<int32_t>(), X.s[2].as<int32_t>(), X.s[3].as<int32_t>(), X.s[4].as<int32_t>(),
                 X.s[5].as<int32_t>(), X.s[6].as<int32_t>()};
        ExtractParams P;
        P.min_size = c->P.min_size; P.max_size = c->P.max_size; P.min_mapq = c->P.min_mapq; P.max_split_parts = c->P.max_split_parts;
        P.min_read_len = c->P.min_read_len; P.min_siglength = c->P.min_siglength; P.merge_del_threshold = c->P.merge_del_threshold;
        P.merge_ins_threshold = c->P.merge_ins_threshold;
        stage_begin(c, CSV_ST_EXTRACT);
        if (n > 0) {
            int per_sm = 1;
            CU(cudaFuncSetAttribute(k_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, EX_SMEM_BYTES));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extract, EX_THREADS, EX_SMEM_BYTES));
            LAUNCH(c, k_extract, grid_for(c, n * 32, EX_THREADS, std::max(per_sm, 1)), EX_THREADS, EX_SMEM_BYTES, R, X.cigar.as<uint32_t>(), S, P, O,
                   (int32_t)X.n_records, dc + 9);
        }
        stage_end(c, CSV_ST_EXTRACT);
        CU(cudaMemcpyAsync(X.h_counters, dc, 16 * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint32_t* h = X.h_counters;
        bool over = false;
        for (int t = 0; t < CSV_NTYPES; t++) if (h[t] > cap[t]) { cap[t] = h[t] + 16; over = true; }
        if (h[5] > cap_pieces) { cap_pieces = h[5] + 16; over = true; }
        if (h[6] > cap_rows) { cap_rows = h[6] + 16; over = true; }
        if (over) continue;
        if (h[7] & ST_INTERNAL) return set_err(CSV_E_CUDA, "csv_extract: internal error (status 0x%x)", h[7]);
        for (int t = 0; t < CSV_NTYPES; t++) {
            c->sig[t].n = h[t];
            c->sig[t].has_c = (t == CSV_INS || t == CSV_INV || t == CSV_TRA);
            if (counts) counts[t] = h[t];
        }
        if (n > 0) {
            for (int t = 0; t < CSV_NTYPES; t++) X.per_record[t] = std::max(X.per_record[t], (double)(h[t] - base[t]) / (double)n);
            X.per_record[5] = std::max(X.per_record[5], (double)(h[5] - base[5]) / (double)n);
        }
        X.n_pieces = h[5];
        X.n_skipped += h[8];
        X.n_records += (uint32_t)n;
        c->n_reads = h[6];
        if (n_read_rows) *n_read_rows = h[6];
        c->counts_valid = false;
        CU(cudaEventRecord(c->ev_done, c->stream));
        c->done_pending = true;
        if (c->profiling) stage_collect(c);
        return CSV_OK;
    }
    return set_err(CSV_E_CUDA, "csv_extract: output capacity did not converge");
}

extern "C" int csv_extract(csv_ctx* c, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar, const csv_sa_cols* sa,
                           int64_t counts[CSV_NTYPES], int64_t* n_read_rows) {
    return extract_impl(c, reads, cigar, n_cigar, sa, counts, n_read_rows, false);
}
extern "C" int csv_extract_append(csv_ctx* c, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar, const csv_sa_cols* sa,
                                  int64_t counts[CSV_NTYPES], int64_t* n_read_rows) {
    return extract_impl(c, reads, cigar, n_cigar, sa, counts, n_read_rows, true);
}
extern "C" int csv_extract_reset(csv_ctx* c) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    for (int t = 0; t < CSV_NTYPES; t++) c->sig[t].n = 0;
    c->n_reads = 0;
    c->ex.n_pieces = 0; c->ex.n_records = 0; c->ex.n_skipped = 0; c->ex.appending = false;
    c->counts_valid = false;
    return CSV_OK;
}
extern "C" int64_t csv_extract_skipped(csv_ctx* c) { return c ? (int64_t)c->ex.n_skipped : 0; }

extern "C" int csv_remap_read_ids(csv_ctx* c, const int32_t* rank, int64_t n_rank) {
    if (!c || !rank || n_rank < 0) return set_err(CSV_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(c->cal_in0.ensure((size_t)std::max<int64_t>(n_rank, 1) * 4));
    CU(c->aln_flag.ensure(64));
    CU(cudaMemsetAsync(c->aln_flag.p, 0, 4, c->stream));
    CU(cudaMemcpyAsync(c->cal_in0.p, rank, (size_t)n_rank * 4, cudaMemcpyHostToDevice, c->stream));
    for (int t = 0; t < CSV_NTYPES; t++)
        if (c->sig[t].n > 0)
            LAUNCH(c, k_remap_ids, grid_for(c, c->sig[t].n, 256), 256, 0, c->sig[t].rid.as<int32_t>(), c->sig[t].n, c->cal_in0.as<int32_t>(), n_rank,
                   c->aln_flag.as<uint32_t>());
    if (c->n_reads > 0)
        LAUNCH(c, k_remap_ids, grid_for(c, c->n_reads, 256), 256, 0, c->r_id.as<int32_t>(), c->n_reads, c->cal_in0.as<int32_t>(), n_rank,
               c->aln_flag.as<uint32_t>());
    uint32_t st = 0;
    CU(cudaMemcpyAsync(&st, c->aln_flag.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->counts_valid = false;
    if (st) return set_err(CSV_E_INPUT, "csv_remap_read_ids: a read id is outside the rank table");
    return CSV_OK;
}

// rows i <-> j of the device-resident INS signatures (and their piece descriptors): the host orders INS rows that tie on
// (contig, int(pos), len, read) by their sequence strings, as the reference's sort key does (cuteSV:774)
__global__ void k_swap_rows(int32_t* c0, int32_t* c1, int32_t* c2, int32_t* c3, int32_t* c4, int32_t* c5, int32_t* c6, const int64_t* pairs, int64_t n_pairs) {
    int32_t* cols[7] = {c0, c1, c2, c3, c4, c5, c6};
    if (blockIdx.x == 0 && threadIdx.x == 0)   // sequential: a row may take part in several swaps
        for (int64_t p = 0; p < n_pairs; p++) {
            const int64_t i = pairs[2 * p], j = pairs[2 * p + 1];
            for (int k = 0; k < 7; k++) if (cols[k]) { const int32_t x = cols[k][i]; cols[k][i] = cols[k][j]; cols[k][j] = x; }
        }
}
extern "C" int csv_swap_ins_rows(csv_ctx* c, const int64_t* pairs, int64_t n_pairs) {
    if (!c || (!pairs && n_pairs) || n_pairs < 0) return set_err(CSV_E_INVALID, "bad argument");
    if (n_pairs == 0) return CSV_OK;
    CU(cudaSetDevice(c->device));
    SigBuf& s = c->sig[CSV_INS];
    for (int64_t p = 0; p < 2 * n_pairs; p++) if (pairs[p] < 0 || pairs[p] >= s.n) return set_err(CSV_E_INVALID, "row index out of range");
    CU(c->cal_out.ensure((size_t)n_pairs * 16));
    CU(cudaMemcpyAsync(c->cal_out.p, pairs, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, c->stream));
    const bool ex = c->ex.piece_off.p && c->ex.piece_off.cap >= (size_t)s.n * 4;
    LAUNCH(c, k_swap_rows, 1, 32, 0, s.chrom.as<int32_t>(), s.a.as<int32_t>(), s.b.as<int32_t>(), s.rid.as<int32_t>(), s.c.as<int32_t>(),
           ex ? c->ex.piece_off.as<int32_t>() : (int32_t*)nullptr, ex ? c->ex.piece_cnt.as<int32_t>() : (int32_t*)nullptr,
           c->cal_out.as<int64_t>(), n_pairs);
    CU(cudaStreamSynchronize(c->stream));
    c->counts_valid = false;
    return CSV_OK;
}

static int fetch_col(csv_ctx* c, void* dst, const DBuf& src, size_t bytes) {
    if (dst && bytes) CU(cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    return CSV_OK;
}

static int fetch_sigs_range(csv_ctx* c, int t, int64_t first, int64_t count, int32_t* chrom, int32_t* a, int32_t* b, int32_t* read_id, int32_t* cc,
                            int32_t* piece_off, int32_t* piece_cnt) {
    const SigBuf& s = c->sig[t];
    const size_t bytes = (size_t)count * 4, o = (size_t)first * 4;
    auto col = [&](void* dst, const DBuf& src) -> int {
        if (dst && bytes) CU(cudaMemcpyAsync(dst, (const char*)src.p + o, bytes, cudaMemcpyDeviceToHost, c->stream));
        return CSV_OK;
    };
    int rc;
    if ((rc = col(chrom, s.chrom)) || (rc = col(a, s.a)) || (rc = col(b, s.b)) || (rc = col(read_id, s.rid))) return rc;
    if (s.has_c && (rc = col(cc, s.c))) return rc;
    if (t == CSV_INS && c->ex.piece_off.p && ((rc = col(piece_off, c->ex.piece_off)) || (rc = col(piece_cnt, c->ex.piece_cnt)))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return CSV_OK;
}
extern "C" int csv_fetch_sigs(csv_ctx* c, int t, int64_t cap, int32_t* chrom, int32_t* a, int32_t* b, int32_t* read_id, int32_t* cc,
                              int32_t* piece_off, int32_t* piece_cnt) {
    if (!c || t < 0 || t >= CSV_NTYPES) return set_err(CSV_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    if (c->sig[t].n > cap) return set_err(CSV_E_CAPACITY, "need %lld", (long long)c->sig[t].n);
    return fetch_sigs_range(c, t, 0, c->sig[t].n, chrom, a, b, read_id, cc, piece_off, piece_cnt);
}
extern "C" int csv_fetch_sigs_range(csv_ctx* c, int t, int64_t first, int64_t count, int32_t* chrom, int32_t* a, int32_t* b, int32_t* read_id,
                                    int32_t* cc, int32_t* piece_off, int32_t* piece_cnt) {
    if (!c || t < 0 || t >= CSV_NTYPES || first < 0 || count < 0) return set_err(CSV_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    if (first + count > c->sig[t].n) return set_err(CSV_E_INVALID, "range [%lld, %lld) outside the %lld signatures", (long long)first, (long long)(first + count), (long long)c->sig[t].n);
    return fetch_sigs_range(c, t, first, count, chrom, a, b, read_id, cc, piece_off, piece_cnt);
}
extern "C" int csv_fetch_pieces_range(csv_ctx* c, int64_t first, int64_t count, int32_t* pieces4) {
    if (!c || first < 0 || count < 0 || !pieces4) return set_err(CSV_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    if (first + count > (int64_t)c->ex.n_pieces) return set_err(CSV_E_INVALID, "piece range outside the table");
    if (count) CU(cudaMemcpyAsync(pieces4, (const char*)c->ex.pieces.p + (size_t)first * sizeof(InsPiece), (size_t)count * sizeof(InsPiece), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CSV_OK;
}

extern "C" int csv_fetch_pieces(csv_ctx* c, int64_t cap, int32_t* pieces4, int64_t* n_pieces) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    CU(cudaSetDevice(c->device));
    if (n_pieces) *n_pieces = c->ex.n_pieces;
    if (!pieces4) return CSV_OK;
    if ((int64_t)c->ex.n_pieces > cap) return set_err(CSV_E_CAPACITY, "need %u", c->ex.n_pieces);
    if (c->ex.n_pieces) CU(cudaMemcpyAsync(pieces4, c->ex.pieces.p, (size_t)c->ex.n_pieces * sizeof(InsPiece), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CSV_OK;
}

extern "C" int csv_fetch_read_rows(csv_ctx* c, int64_t cap, int32_t* chrom, int32_t* start, int32_t* end, int32_t* read_id,
                                   uint8_t* is_primary) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    CU(cudaSetDevice(c->device));
    if (c->n_reads > cap) return set_err(CSV_E_CAPACITY, "need %lld", (long long)c->n_reads);
    const size_t bytes = (size_t)c->n_reads * 4;
    int rc;
    if ((rc = fetch_col(c, chrom, c->r_chrom, bytes))) return rc;
    if ((rc = fetch_col(c, start, c->r_start, bytes))) return rc;
    if ((rc = fetch_col(c, end, c->r_end, bytes))) return rc;
    if ((rc = fetch_col(c, read_id, c->r_id, bytes))) return rc;
    if ((rc = fetch_col(c, is_primary, c->r_prim, (size_t)c->n_reads))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return CSV_OK;
}
