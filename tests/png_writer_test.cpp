// tests/png_writer_test.cpp — RenderBuffer::save_to_png of the host layer on a known buffer
// (no GPU involved): the test decodes the file and expects the reference's conversion
// (render_buffer.h:35-55: row 0 of the buffer is the BOTTOM row of the image, (unsigned char)(x * 255)).
#include "rtb_host.hpp"

int main(int argc, char **argv) {
    if (argc < 4)
        return 2;
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    RenderBuffer buf(w, h);
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) // values cover 0, 1 and fractions on both sides of a byte boundary
            buf.set_pixel(i, j, color((i + 0.5) / w, (j == h - 1) ? 1.0 : double(j) / h, ((i * 7 + j * 13) % 256) / 255.0 + 1e-9));
    buf.set_pixel(-1, 0, color(9, 9, 9)); // out of range: ignored (render_buffer.h:18-22)
    buf.set_pixel(0, h, color(9, 9, 9));
    return buf.save_to_png(argv[1]) ? 0 : 1;
}
