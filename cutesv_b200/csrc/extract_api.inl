// extract_api.inl -- C-ABI of the extraction stage (included by cutesv_b200.cu)
extern "C" int csv_extract(csv_ctx* c, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar, const csv_sa_cols* sa,
                           int64_t counts[CSV_NTYPES], int64_t* n_read_rows) {
    (void)c; (void)reads; (void)cigar; (void)n_cigar; (void)sa; (void)counts; (void)n_read_rows;
    return set_err(CSV_E_STATE, "csv_extract: not built in this revision");
}
extern "C" int csv_fetch_sigs(csv_ctx* c, int svtype, int64_t cap, int32_t* chrom, int32_t* a, int32_t* b, int32_t* read_id, int32_t* cc,
                              int32_t* extra3) {
    (void)c; (void)svtype; (void)cap; (void)chrom; (void)a; (void)b; (void)read_id; (void)cc; (void)extra3;
    return set_err(CSV_E_STATE, "csv_fetch_sigs: not built in this revision");
}
extern "C" int csv_fetch_read_rows(csv_ctx* c, int64_t cap, int32_t* chrom, int32_t* start, int32_t* end, int32_t* read_id,
                                   uint8_t* is_primary) {
    (void)c; (void)cap; (void)chrom; (void)start; (void)end; (void)read_id; (void)is_primary;
    return set_err(CSV_E_STATE, "csv_fetch_read_rows: not built in this revision");
}
