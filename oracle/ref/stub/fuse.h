/*
 * Minimal stand-in for libfuse 2.6's <fuse.h>, written for the oracle build only.
 * TEST INFRASTRUCTURE: lets the UNMODIFIED reference sources under /root/reference
 * compile in a container without libfuse.  Nothing here mounts anything; the harness
 * (ref_harness.c) drives the reference's process_frame() directly.
 */
#ifndef ORACLE_STUB_FUSE_H
#define ORACLE_STUB_FUSE_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>
#include <sys/stat.h>
#include <sys/statvfs.h>

struct fuse_file_info { int flags; uint64_t fh; };

typedef int (*fuse_fill_dir_t)(void *buf, const char *name, const struct stat *stbuf, off_t off);

struct fuse_operations {
    int (*getattr)(const char *, struct stat *);
    int (*open)(const char *, struct fuse_file_info *);
    int (*read)(const char *, char *, size_t, off_t, struct fuse_file_info *);
    int (*readdir)(const char *, void *, fuse_fill_dir_t, off_t, struct fuse_file_info *);
    int (*create)(const char *, mode_t, struct fuse_file_info *);
    int (*fsync)(const char *, int, struct fuse_file_info *);
    int (*mkdir)(const char *, mode_t);
    int (*release)(const char *, struct fuse_file_info *);
    int (*rename)(const char *, const char *);
    int (*rmdir)(const char *);
    int (*truncate)(const char *, off_t);
    int (*write)(const char *, const char *, size_t, off_t, struct fuse_file_info *);
    int (*statfs)(const char *, struct statvfs *);
    int (*unlink)(const char *);
};

struct fuse_opt { const char *templ; unsigned long offset; int value; };
struct fuse_args { int argc; char **argv; int allocated; };

#define FUSE_ARGS_INIT(c, v) { (c), (v), 0 }
#define FUSE_OPT_END { NULL, 0, 0 }

typedef int (*fuse_opt_proc_t)(void *data, const char *arg, int key, struct fuse_args *outargs);

int fuse_opt_parse(struct fuse_args *args, void *data, const struct fuse_opt opts[], fuse_opt_proc_t proc);
void fuse_opt_free_args(struct fuse_args *args);
int fuse_main(int argc, char *argv[], const struct fuse_operations *op, void *user_data);

#endif
