/*
 * oracle/ref/ref_harness.c -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * Wraps the UNMODIFIED reference (compiled where it lies under /root/reference, see
 * oracle/Makefile) so that tests and bench.py's cpu_baseline / --impl reference legs can
 * drive the reference's own frame builder:
 *
 *   - process_frame() is `static` in mlvfs/main.c:908, so this translation unit includes
 *     main.c textually (with main() renamed) and re-exports a thin caller.
 *   - struct mlvfs (mlvfs/mlvfs.h:32-48) is the file-static option block main.c:53; the
 *     setters below write its fields the way fuse_opt_parse / webgui.c:298-336 would.
 *
 * Every other reference entry point (dng_get_image_data, chroma_smooth, fix_bad_pixels,
 * fix_focus_pixels, stripes_*, fix_pattern_noise, hdr_convert_data, cr2hdr20_convert_data,
 * lj92_*, get_raw2ev, get_ev2raw ...) is a normal exported symbol of the resulting
 * libmlvfs_ref.so and is called directly through ctypes.
 */
#define main mlvfs_reference_main
#include "main.c"
#undef main

#include <stddef.h>

/* libfuse entry points referenced by main.c; never reached by the harness */
int fuse_opt_parse(struct fuse_args *args, void *data, const struct fuse_opt opts[], fuse_opt_proc_t proc)
{ (void)args; (void)data; (void)opts; (void)proc; return 0; }
void fuse_opt_free_args(struct fuse_args *args) { (void)args; }
int fuse_main(int argc, char *argv[], const struct fuse_operations *op, void *user_data)
{ (void)argc; (void)argv; (void)op; (void)user_data; return 0; }

/* one-time LUT construction, same order as main.c:1948-1950 */
void ref_init(void)
{
    get_raw2evf(0);
    get_raw2ev(0);
    get_ev2raw();
}

/* option block: fields of struct mlvfs, mlvfs.h:32-48 */
void ref_set_mlv_dir(const char *dir)
{
    static char dirbuf[4096];
    strncpy(dirbuf, dir, sizeof(dirbuf) - 1);
    mlvfs.mlv_path = dirbuf;
}

void ref_set_options(int chroma_smooth_, int fix_bad_pixels_, int fix_stripes_, int dual_iso_,
                     int hdr_interpolation_method_, int hdr_no_fullres_, int hdr_no_alias_map_,
                     int fix_pattern_noise_, int deflicker_)
{
    mlvfs.chroma_smooth = chroma_smooth_;
    mlvfs.fix_bad_pixels = fix_bad_pixels_;
    mlvfs.fix_stripes = fix_stripes_;
    mlvfs.dual_iso = dual_iso_;
    mlvfs.hdr_interpolation_method = hdr_interpolation_method_;
    mlvfs.hdr_no_fullres = hdr_no_fullres_;
    mlvfs.hdr_no_alias_map = hdr_no_alias_map_;
    mlvfs.fix_pattern_noise = fix_pattern_noise_;
    mlvfs.deflicker = deflicker_;
}

/*
 * Run the reference's process_frame (main.c:908-1005) for "/<clip>.MLV/<clip>_NNNNNN.dng".
 * Copies image_buffer.data into out (cap bytes) and, if header_out != NULL, the 64 KiB DNG
 * header.  Returns the image size in bytes, 0 on failure.
 */
size_t ref_process_frame(const char *virtual_dng_path, uint16_t *out, size_t cap, uint8_t *header_out)
{
    struct image_buffer ib;
    memset(&ib, 0, sizeof(ib));
    ib.dng_filename = (char *)virtual_dng_path;
    process_frame(&ib);
    if (!ib.data) return 0;
    size_t n = ib.size < cap ? ib.size : cap;
    if (out) memcpy(out, ib.data, n);
    if (header_out && ib.header) memcpy(header_out, ib.header, ib.header_size);
    size_t sz = ib.size;
    free(ib.data);
    free(ib.header);
    return sz;
}

/* the reference's own header walk (main.c:429-558), used to pin our index reader */
int ref_get_frame_headers(const char *mlv_path, int index, struct frame_headers *out)
{
    return mlv_get_frame_headers(mlv_path, index, out);
}

int ref_get_frame_count(const char *mlv_path)
{
    return mlv_get_frame_count(mlv_path);
}

/* layout probes: pin include/mlvb_mlv_format.h against the reference's packed structs */
size_t ref_sizeof_frame_headers(void) { return sizeof(struct frame_headers); }
size_t ref_offsetof_frame_headers(int field)
{
    switch (field) {
    case 0: return offsetof(struct frame_headers, fileNumber);
    case 1: return offsetof(struct frame_headers, position);
    case 2: return offsetof(struct frame_headers, vidf_hdr);
    case 3: return offsetof(struct frame_headers, file_hdr);
    case 4: return offsetof(struct frame_headers, rtci_hdr);
    case 5: return offsetof(struct frame_headers, idnt_hdr);
    case 6: return offsetof(struct frame_headers, rawi_hdr);
    case 7: return offsetof(struct frame_headers, expo_hdr);
    case 8: return offsetof(struct frame_headers, lens_hdr);
    case 9: return offsetof(struct frame_headers, wbal_hdr);
    case 10: return offsetof(struct frame_headers, rawi_hdr) + offsetof(mlv_rawi_hdr_t, raw_info);
    case 11: return offsetof(struct raw_info, black_level);
    case 12: return offsetof(struct raw_info, white_level);
    case 13: return offsetof(struct raw_info, bits_per_pixel);
    case 14: return offsetof(struct raw_info, frame_size);
    case 15: return sizeof(struct raw_info);
    case 16: return offsetof(struct raw_info, exposure_bias);
    case 17: return offsetof(struct raw_info, dng_active_area);
    default: return (size_t)-1;
    }
}

/* main.c:895-906 (static): per-frame exposure compensation written into raw_info.exposure_bias */
void ref_deflicker(struct frame_headers *hdrs, int target, uint16_t *data, size_t size_bytes)
{
    deflicker(hdrs, target, data, size_bytes);
}
