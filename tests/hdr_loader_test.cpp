// tests/hdr_loader_test.cpp — EnvironmentLight(const char*) of the host layer on a Radiance .hdr
// file (no GPU involved): dumps width, height and the float texels so the test can compare them
// with what the reference's stbi_loadf returns for the same file (environmental_light.h:121-134).
#include "rtb_host.hpp"

int main(int argc, char **argv) {
    if (argc < 3)
        return 2;
    EnvironmentLight env(argv[1]);
    std::ofstream out(argv[2], std::ios::binary);
    const int32_t dims[2] = {env.width, env.height};
    out.write(reinterpret_cast<const char *>(dims), sizeof(dims));
    out.write(reinterpret_cast<const char *>(env.hdr_data.data()), std::streamsize(env.hdr_data.size() * sizeof(float)));
    return out ? 0 : 1;
}
