/*
 * cnngp_h5.h -- C ABI of the native HDF5 block store (libcnngp_h5.so, host code only).
 *
 * The reference keeps Gram matrices in HDF5 files through h5py:
 *     h5py.File(path, "w" | "a" | "r")                       exp_mnist_resnet/save_kernel.py:26,33
 *     f.create_dataset(name, shape=(1,N,N2), dtype=float32,  cnn_gp/kernel_save_tools.py:21-23
 *                      fillvalue=nan, chunks=(1,bs,bs), maxshape=(None,N,N2))
 *     out[0, i:i+n, j:j+m] = k                               cnn_gp/kernel_save_tools.py:55-58
 *     dset.read_direct(A, source_sel=np.s_[i, :, :])         exp_mnist_resnet/classify_gp.py:45-48
 *     dest[isnan(dest)] = src[isnan(dest)]                   exp_mnist_resnet/merge_h5_files.py:24-30
 * h5py / libhdf5 are not part of the B200 image, and this file format is the wire contract
 * between save_kernel, merge_h5_files and classify_gp.  This library writes and reads that
 * format itself -- HDF5 File Format Specification, the structures libhdf5's default
 * ("earliest") settings produce: version-0 superblock, symbol-table root group (version-1
 * B-tree, SNOD, local heap), version-1 object headers, chunked storage indexed by a version-1
 * B-tree, IEEE little-endian float32 / float64 elements, fill-value message -- so files
 * written here open in h5py / h5dump, and files h5py wrote with default settings (unfiltered,
 * flat namespace) open here.
 *
 * Conventions: plain C types; 0 = success, non-zero = error with a thread-local message in
 * cnngp_h5_last_error(); all buffers are HOST memory owned by the caller, C-contiguous, in the
 * dataset's element type; a handle may be used from several threads (calls serialise on it).
 * Selections are unit-stride hyperslabs (start[], count[]), which is all the path uses.
 */
#ifndef CNNGP_H5_H
#define CNNGP_H5_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNNGP_H5_MAX_RANK 8
#define CNNGP_H5_UNLIMITED (-1) /* maxshape entry: unlimited (h5py: None) */

typedef struct cnngp_h5 cnngp_h5; /* an open file */

const char *cnngp_h5_last_error(void);

/* mode: "r" read-only | "r+" read/write, must exist | "w" create/truncate | "w-" or "x" create,
 * fail if it exists | "a" read/write, create if missing          (h5py.File's modes) */
int cnngp_h5_open(const char *path, const char *mode, cnngp_h5 **out);
/* write every dirty index / header and the end-of-file address; the file is a valid HDF5 file
 * after each flush */
int cnngp_h5_flush(cnngp_h5 *f);
int cnngp_h5_close(cnngp_h5 *f); /* flush + release; the handle is gone afterwards even on error */

/* datasets of the root group, in name order (f.keys()) */
int cnngp_h5_count(cnngp_h5 *f);
int cnngp_h5_name(cnngp_h5 *f, int index, char *buf, int cap); /* returns the length needed incl. NUL */
int cnngp_h5_find(cnngp_h5 *f, const char *name);               /* dataset id >= 0, or -1 */

/* f.create_dataset(name, shape, dtype, fillvalue=, chunks=, maxshape=).
 *   dtype    0 = float32, 1 = float64 (IEEE, little-endian)
 *   chunks   NULL = contiguous storage (maxshape must then equal shape), else the chunk shape
 *   maxshape NULL = shape; CNNGP_H5_UNLIMITED entries allowed for chunked datasets
 *   fill     NULL = no fill value defined (reads of unwritten elements give 0), else one element
 * Returns the dataset id through *id. */
int cnngp_h5_create_dataset(cnngp_h5 *f, const char *name, int rank, const int64_t *shape,
                            const int64_t *maxshape, const int64_t *chunks, int dtype, const void *fill,
                            int *id);

typedef struct cnngp_h5_info {
    int32_t rank, dtype;         /* dtype -1: an element type this library does not read */
    int32_t chunked, has_fill;
    int64_t shape[CNNGP_H5_MAX_RANK], maxshape[CNNGP_H5_MAX_RANK], chunks[CNNGP_H5_MAX_RANK];
    double fill;                 /* the fill value widened to double */
    int64_t n_chunks_stored;     /* chunks that exist in the file (0 for contiguous) */
} cnngp_h5_info;
int cnngp_h5_dataset_info(cnngp_h5 *f, int id, cnngp_h5_info *info);

/* dset[start : start+count] = data  /  data = dset[start : start+count]  (rank entries each).
 * Elements never written read as the fill value. */
int cnngp_h5_write(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, const void *data);
int cnngp_h5_read(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, void *data);
/* The same for a caller's array that is a strided view (a column range of a wider row buffer,
 * as save_K_resident hands over): stride_bytes[i] = distance between consecutive indices of
 * dimension i; the last dimension must be contiguous (stride = element size).  Saves the packing
 * copy h5py makes for such selections. */
int cnngp_h5_write_strided(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, const void *data,
                           const int64_t *stride_bytes);
int cnngp_h5_read_strided(cnngp_h5 *f, int id, const int64_t *start, const int64_t *count, void *data,
                          const int64_t *stride_bytes);
/* dset.resize(new_shape): every extent within maxshape; stored chunks outside the new extent
 * stay in the file but are no longer reachable */
int cnngp_h5_resize(cnngp_h5 *f, int id, const int64_t *new_shape);

/* exp_mnist_resnet/merge_h5_files.py:24-30 for one dataset, chunk by chunk: wherever dest is
 * NaN take src.  Both datasets must have the same shape, element type and chunk shape.  Chunks
 * src never stored are skipped, chunks dest never stored are copied whole -- the cost is
 * proportional to what the workers actually wrote, not to N x N2. */
int cnngp_h5_merge_nan(cnngp_h5 *dest, int dest_id, cnngp_h5 *src, int src_id);

#ifdef __cplusplus
}
#endif
#endif /* CNNGP_H5_H */
