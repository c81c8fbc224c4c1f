/* TEST INFRASTRUCTURE (oracle/): see ref_png_shim.h. */
#include "mve/image_io.h"
namespace mve { namespace image {
void save_png_file (ByteImage::ConstPtr image, std::string const& filename, int)
{
    save_mvei_file(image, filename);
}
} }
