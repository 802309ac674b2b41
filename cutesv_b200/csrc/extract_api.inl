// extract_api.inl -- C-ABI of the extraction stage (included by cutesv_b200.cu)

static int ex_upload(csv_ctx* c, DBuf& b, const void* src, size_t bytes) {
    CU(b.ensure(bytes ? bytes : 4));
    if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return CSV_OK;
}

extern "C" int csv_extract(csv_ctx* c, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar, const csv_sa_cols* sa,
                           int64_t counts[CSV_NTYPES], int64_t* n_read_rows) {
    if (!c || !reads) return set_err(CSV_E_INVALID, "null argument");
    if (c->n_contigs == 0) return set_err(CSV_E_STATE, "csv_set_contigs has not been called");
    const int64_t n = reads->n;
    if (n < 0 || n >= (1ll << 30)) return set_err(CSV_E_INVALID, "record count out of range");
    CU(cudaSetDevice(c->device));
    ExtractState& X = c->ex;
    const int64_t n_sa = sa ? sa->n : 0;
    for (int slot = 0; slot <= CSV_NTYPES; slot++) {  // extraction overwrites the device-resident inputs: drain pending uploads first
        int wrc = wait_upload(c, slot);
        if (wrc) return wrc;
    }
    stage_begin(c, CSV_ST_H2D);
    int rc;
    const void* rsrc[7] = {reads->chrom, reads->ref_start, reads->ref_end, reads->flag, reads->mapq, reads->query_len, reads->read_id};
    for (int k = 0; k < 7; k++) { rc = ex_upload(c, X.r[k], rsrc[k], (size_t)n * 4); if (rc) return rc; }
    rc = ex_upload(c, X.cigar_off, reads->cigar_off, (size_t)(n + 1) * 8); if (rc) return rc;
    rc = ex_upload(c, X.sa_off, reads->sa_off, (size_t)(n + 1) * 8); if (rc) return rc;
    rc = ex_upload(c, X.cigar, cigar, (size_t)n_cigar * 4); if (rc) return rc;
    const void* ssrc[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (n_sa) { ssrc[0] = sa->chrom; ssrc[1] = sa->pos0; ssrc[2] = sa->strand; ssrc[3] = sa->mapq; ssrc[4] = sa->first_clip; ssrc[5] = sa->last_clip; ssrc[6] = sa->ref_span; }
    for (int k = 0; k < 7; k++) { rc = ex_upload(c, X.s[k], ssrc[k], (size_t)n_sa * 4); if (rc) return rc; }
    stage_end(c, CSV_ST_H2D);
    CU(X.counters.ensure(16 * 4));
    if (!X.h_counters) CU(cudaMallocHost((void**)&X.h_counters, 16 * 4));
    // first guess of the output capacities; the kernel keeps counting past them, so one rerun suffices
    uint32_t cap[CSV_NTYPES], cap_pieces, cap_rows = (uint32_t)n + 16;
    cap[CSV_DEL] = cap[CSV_INS] = (uint32_t)std::min<int64_t>(4 * n + 1024, (1ll << 30) - 1);
    cap[CSV_INV] = cap[CSV_DUP] = cap[CSV_TRA] = (uint32_t)std::min<int64_t>(2 * n_sa + n / 4 + 1024, (1ll << 30) - 1);
    cap_pieces = cap[CSV_INS] * 2;
    for (int attempt = 0; attempt < 3; attempt++) {
        ExtractOut O;
        memset(&O, 0, sizeof(O));
        for (int t = 0; t < CSV_NTYPES; t++) {
            SigBuf& sb = c->sig[t];
            CU(sb.chrom.ensure((size_t)cap[t] * 4)); CU(sb.a.ensure((size_t)cap[t] * 4)); CU(sb.b.ensure((size_t)cap[t] * 4));
            CU(sb.rid.ensure((size_t)cap[t] * 4)); CU(sb.c.ensure((size_t)cap[t] * 4));
            O.col[t][0] = sb.chrom.as<int32_t>(); O.col[t][1] = sb.a.as<int32_t>(); O.col[t][2] = sb.b.as<int32_t>();
            O.col[t][3] = sb.rid.as<int32_t>(); O.col[t][4] = sb.c.as<int32_t>();
            O.cap_sig[t] = cap[t];
        }
        CU(X.piece_off.ensure((size_t)cap[CSV_INS] * 4)); CU(X.piece_cnt.ensure((size_t)cap[CSV_INS] * 4));
        CU(X.pieces.ensure((size_t)cap_pieces * sizeof(InsPiece)));
        CU(c->r_chrom.ensure((size_t)cap_rows * 4)); CU(c->r_start.ensure((size_t)cap_rows * 4)); CU(c->r_end.ensure((size_t)cap_rows * 4));
        CU(c->r_id.ensure((size_t)cap_rows * 4)); CU(c->r_prim.ensure((size_t)cap_rows));
        uint32_t* dc = X.counters.as<uint32_t>();
        O.n_sig = dc; O.n_pieces = dc + 5; O.n_rows = dc + 6; O.status = dc + 7;
        O.ins_piece_off = X.piece_off.as<int32_t>(); O.ins_piece_cnt = X.piece_cnt.as<int32_t>(); O.pieces = X.pieces.as<InsPiece>();
        O.cap_pieces = cap_pieces;
        O.rr_chrom = c->r_chrom.as<int32_t>(); O.rr_start = c->r_start.as<int32_t>(); O.rr_end = c->r_end.as<int32_t>();
        O.rr_id = c->r_id.as<int32_t>(); O.rr_prim = c->r_prim.as<uint8_t>(); O.cap_rows = cap_rows;
        CU(cudaMemsetAsync(dc, 0, 16 * 4, c->stream));
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
        if (n > 0) LAUNCH(c, k_extract, grid_for(c, n * 32, EX_THREADS, 8), EX_THREADS, 0, R, X.cigar.as<uint32_t>(), S, P, O, 0);
        stage_end(c, CSV_ST_EXTRACT);
        CU(cudaMemcpyAsync(X.h_counters, dc, 16 * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint32_t* h = X.h_counters;
        bool over = false;
        for (int t = 0; t < CSV_NTYPES; t++) if (h[t] > cap[t]) { cap[t] = h[t] + 16; over = true; }
        if (h[5] > cap_pieces) { cap_pieces = h[5] + 16; over = true; }
        if (h[6] > cap_rows) { cap_rows = h[6] + 16; over = true; }
        if (over) continue;
        if (h[7] & ST_INTERNAL) return set_err(CSV_E_INPUT, "csv_extract: a read exceeds the segment / merged-piece limits (%d / %d)", MAX_SEGS, MAX_OPEN_PIECES);
        for (int t = 0; t < CSV_NTYPES; t++) {
            c->sig[t].n = h[t];
            c->sig[t].has_c = (t == CSV_INS || t == CSV_INV || t == CSV_TRA);
            if (counts) counts[t] = h[t];
        }
        X.n_pieces = h[5];
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

static int fetch_col(csv_ctx* c, void* dst, const DBuf& src, size_t bytes) {
    if (dst && bytes) CU(cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    return CSV_OK;
}

extern "C" int csv_fetch_sigs(csv_ctx* c, int t, int64_t cap, int32_t* chrom, int32_t* a, int32_t* b, int32_t* read_id, int32_t* cc,
                              int32_t* piece_off, int32_t* piece_cnt) {
    if (!c || t < 0 || t >= CSV_NTYPES) return set_err(CSV_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    const SigBuf& s = c->sig[t];
    if (s.n > cap) return set_err(CSV_E_CAPACITY, "need %lld", (long long)s.n);
    const size_t bytes = (size_t)s.n * 4;
    int rc;
    if ((rc = fetch_col(c, chrom, s.chrom, bytes))) return rc;
    if ((rc = fetch_col(c, a, s.a, bytes))) return rc;
    if ((rc = fetch_col(c, b, s.b, bytes))) return rc;
    if ((rc = fetch_col(c, read_id, s.rid, bytes))) return rc;
    if (s.has_c && (rc = fetch_col(c, cc, s.c, bytes))) return rc;
    if (t == CSV_INS && c->ex.piece_off.p) {
        if ((rc = fetch_col(c, piece_off, c->ex.piece_off, bytes))) return rc;
        if ((rc = fetch_col(c, piece_cnt, c->ex.piece_cnt, bytes))) return rc;
    }
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
