// dualiso.cu -- dual-ISO entry points (reference hdr.c).  NOT BUILT YET in this round: both entry
// points report failure loudly (return 0 = "not converted", exactly what the reference returns when it
// cannot convert a frame, hdr.c:1953-1956), so a caller never receives silently unprocessed data
// labelled as converted.  mlvb_process_frame returns MLVB_ERR_UNSUPPORTED for dual_iso != 0.
#include "context.cuh"

extern "C" {

int hdr_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, off_t offset, size_t max_size)
{
    (void)frame_headers; (void)image_data; (void)offset; (void)max_size;
    fprintf(stderr, "libmlvfs_b200: hdr_convert_data (dual-ISO preview, hdr.c:40-227) is not implemented yet\n");
    return 0;
}

int cr2hdr20_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, int interp_method, int fullres,
                          int use_alias_map, int chroma_smooth, int fix_bad_pixels_mode)
{
    (void)frame_headers; (void)image_data; (void)interp_method; (void)fullres; (void)use_alias_map;
    (void)chroma_smooth; (void)fix_bad_pixels_mode;
    fprintf(stderr, "libmlvfs_b200: cr2hdr20_convert_data (dual ISO, hdr.c:1932-1957) is not implemented yet\n");
    return 0;
}

}  // extern "C"
