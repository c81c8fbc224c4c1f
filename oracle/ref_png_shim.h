/* TEST INFRASTRUCTURE (oracle/): force-included (-include) when compiling the
 * UNMODIFIED reference file libs/mve/view.cc with -DMVE_NO_PNG_SUPPORT.
 * The reference calls image::save_png_file unguarded at libs/mve/view.cc:854
 * while its declaration (libs/mve/image_io.h:94-121) is compiled out; libpng is
 * not available in this image. This header re-declares the symbol; the
 * definition in ref_png_shim.cc writes MVEI bytes instead (only affects the
 * `undist-L<s>` embedding's file format, never the depth/conf/dz maps). */
#ifndef ORACLE_REF_PNG_SHIM_H
#define ORACLE_REF_PNG_SHIM_H
#ifdef __cplusplus
#include <string>
#include "mve/image.h"
namespace mve { namespace image {
void save_png_file (ByteImage::ConstPtr image, std::string const& filename,
    int compression_level = 1);
} }
#endif
#endif
