// rtb200_blob.hpp — header-only writer / reader for the flat scene blob of
// rtb200_scene.h.  Pure host C++11, no dependencies; used by the host layer
// (scene flattener), by the device library (upload) and by the test oracles.
#ifndef RTB200_BLOB_HPP
#define RTB200_BLOB_HPP

#include "rtb200_scene.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtb {

// Accumulates the tables and serialises them into one contiguous blob.
struct SceneTables {
    rtb_globals globals{};
    rtb_camera camera{};
    std::vector<rtb_prim> prims;
    std::vector<rtb_chain> chains;
    std::vector<rtb_xform_op> xform_ops;
    std::vector<rtb_material> materials;
    std::vector<rtb_texture> textures;
    std::vector<rtb_image> images;
    std::vector<uint8_t> image_bytes;
    std::vector<rtb_perlin> perlins;
    std::vector<rtb_light> lights;
    std::vector<float> env_texels;
    std::vector<rtb_gate> gates;

    std::vector<uint8_t> serialise() const {
        struct Sec {
            uint32_t id, stride;
            uint64_t count;
            const void *data;
        };
        const Sec secs[] = {
            {RTB_SEC_GLOBALS, sizeof(rtb_globals), 1, &globals},
            {RTB_SEC_CAMERA, sizeof(rtb_camera), 1, &camera},
            {RTB_SEC_PRIMS, sizeof(rtb_prim), prims.size(), prims.data()},
            {RTB_SEC_CHAINS, sizeof(rtb_chain), chains.size(), chains.data()},
            {RTB_SEC_XFORM_OPS, sizeof(rtb_xform_op), xform_ops.size(), xform_ops.data()},
            {RTB_SEC_MATERIALS, sizeof(rtb_material), materials.size(), materials.data()},
            {RTB_SEC_TEXTURES, sizeof(rtb_texture), textures.size(), textures.data()},
            {RTB_SEC_IMAGES, sizeof(rtb_image), images.size(), images.data()},
            {RTB_SEC_IMAGE_BYTES, 1, image_bytes.size(), image_bytes.data()},
            {RTB_SEC_PERLIN, sizeof(rtb_perlin), perlins.size(), perlins.data()},
            {RTB_SEC_LIGHTS, sizeof(rtb_light), lights.size(), lights.data()},
            {RTB_SEC_ENV_TEXELS, sizeof(float), env_texels.size(), env_texels.data()},
            {RTB_SEC_GATES, sizeof(rtb_gate), gates.size(), gates.data()},
        };
        const uint32_t n = sizeof(secs) / sizeof(secs[0]);
        uint64_t off = sizeof(rtb_blob_header) + uint64_t(n) * sizeof(rtb_blob_section);
        off = (off + 7) & ~uint64_t(7);
        std::vector<rtb_blob_section> dir(n);
        for (uint32_t i = 0; i < n; ++i) {
            dir[i].id = secs[i].id;
            dir[i].stride = secs[i].stride;
            dir[i].count = secs[i].count;
            dir[i].offset = off;
            off += secs[i].count * secs[i].stride;
            off = (off + 7) & ~uint64_t(7);
        }
        std::vector<uint8_t> blob(off, 0);
        rtb_blob_header h{};
        h.magic = RTB_SCENE_MAGIC;
        h.version = RTB_SCENE_VERSION;
        h.total_bytes = off;
        h.n_sections = n;
        std::memcpy(blob.data(), &h, sizeof(h));
        std::memcpy(blob.data() + sizeof(h), dir.data(), n * sizeof(rtb_blob_section));
        for (uint32_t i = 0; i < n; ++i)
            if (secs[i].count)
                std::memcpy(blob.data() + dir[i].offset, secs[i].data,
                            secs[i].count * secs[i].stride);
        return blob;
    }
};

// Non-owning typed view over a serialised blob.  Throws std::runtime_error on
// a malformed blob (the C-ABI catches and converts to a status code).
class SceneView {
  public:
    SceneView(const void *blob, uint64_t nbytes) : base_(static_cast<const uint8_t *>(blob)) {
        if (!blob || nbytes < sizeof(rtb_blob_header))
            throw std::runtime_error("scene blob: too small");
        std::memcpy(&hdr_, base_, sizeof(hdr_));
        if (hdr_.magic != RTB_SCENE_MAGIC)
            throw std::runtime_error("scene blob: bad magic");
        if (hdr_.version != RTB_SCENE_VERSION)
            throw std::runtime_error("scene blob: unsupported version");
        if (hdr_.total_bytes > nbytes)
            throw std::runtime_error("scene blob: truncated");
        const uint64_t dir_end =
            sizeof(rtb_blob_header) + uint64_t(hdr_.n_sections) * sizeof(rtb_blob_section);
        if (dir_end > hdr_.total_bytes)
            throw std::runtime_error("scene blob: bad section directory");
        dir_.resize(hdr_.n_sections);
        std::memcpy(dir_.data(), base_ + sizeof(rtb_blob_header),
                    hdr_.n_sections * sizeof(rtb_blob_section));
        for (const auto &s : dir_) // (written so that no product or sum can wrap)
            if (s.offset > hdr_.total_bytes || (s.offset & 3u) ||
                (s.count && (s.stride == 0 || s.count > (hdr_.total_bytes - s.offset) / s.stride)))
                throw std::runtime_error("scene blob: section out of range");
        if (count<rtb_globals>(RTB_SEC_GLOBALS) != 1 || count<rtb_camera>(RTB_SEC_CAMERA) != 1)
            throw std::runtime_error("scene blob: missing globals/camera");
    }

    template <typename T> const T *data(uint32_t id) const {
        const rtb_blob_section *s = find(id);
        return s && s->count ? reinterpret_cast<const T *>(base_ + s->offset) : nullptr;
    }
    template <typename T> uint64_t count(uint32_t id) const {
        const rtb_blob_section *s = find(id);
        if (!s)
            return 0;
        if (s->stride != sizeof(T))
            throw std::runtime_error("scene blob: record size mismatch in section " +
                                     std::to_string(id));
        return s->count;
    }

    const rtb_globals &globals() const { return *data<rtb_globals>(RTB_SEC_GLOBALS); }
    const rtb_camera &camera() const { return *data<rtb_camera>(RTB_SEC_CAMERA); }
    const rtb_prim *prims() const { return data<rtb_prim>(RTB_SEC_PRIMS); }
    uint64_t n_prims() const { return count<rtb_prim>(RTB_SEC_PRIMS); }
    const rtb_chain *chains() const { return data<rtb_chain>(RTB_SEC_CHAINS); }
    uint64_t n_chains() const { return count<rtb_chain>(RTB_SEC_CHAINS); }
    const rtb_xform_op *xform_ops() const { return data<rtb_xform_op>(RTB_SEC_XFORM_OPS); }
    uint64_t n_xform_ops() const { return count<rtb_xform_op>(RTB_SEC_XFORM_OPS); }
    const rtb_material *materials() const { return data<rtb_material>(RTB_SEC_MATERIALS); }
    uint64_t n_materials() const { return count<rtb_material>(RTB_SEC_MATERIALS); }
    const rtb_texture *textures() const { return data<rtb_texture>(RTB_SEC_TEXTURES); }
    uint64_t n_textures() const { return count<rtb_texture>(RTB_SEC_TEXTURES); }
    const rtb_image *images() const { return data<rtb_image>(RTB_SEC_IMAGES); }
    uint64_t n_images() const { return count<rtb_image>(RTB_SEC_IMAGES); }
    const uint8_t *image_bytes() const { return data<uint8_t>(RTB_SEC_IMAGE_BYTES); }
    uint64_t n_image_bytes() const { return count<uint8_t>(RTB_SEC_IMAGE_BYTES); }
    const rtb_perlin *perlins() const { return data<rtb_perlin>(RTB_SEC_PERLIN); }
    uint64_t n_perlins() const { return count<rtb_perlin>(RTB_SEC_PERLIN); }
    const rtb_light *lights() const { return data<rtb_light>(RTB_SEC_LIGHTS); }
    uint64_t n_lights() const { return count<rtb_light>(RTB_SEC_LIGHTS); }
    const float *env_texels() const { return data<float>(RTB_SEC_ENV_TEXELS); }
    uint64_t n_env_texels() const { return count<float>(RTB_SEC_ENV_TEXELS); }
    const rtb_gate *gates() const { return data<rtb_gate>(RTB_SEC_GATES); }
    uint64_t n_gates() const { return count<rtb_gate>(RTB_SEC_GATES); }

    // Structural validation of every cross-reference; throws on the first bad one.
    void validate() const {
        const int64_t np = int64_t(n_prims()), nc = int64_t(n_chains()),
                      nx = int64_t(n_xform_ops()), nm = int64_t(n_materials()),
                      nt = int64_t(n_textures()), ni = int64_t(n_images()),
                      npl = int64_t(n_perlins());
        for (int64_t i = 0; i < nc; ++i) {
            const rtb_chain &c = chains()[i];
            if (c.first < 0 || c.count < 0 || int64_t(c.first) + c.count > nx)
                throw std::runtime_error("scene blob: chain out of range");
        }
        for (int64_t i = 0; i < nx; ++i)
            if (xform_ops()[i].kind < 0 || xform_ops()[i].kind > RTB_XF_FLIP_FACE)
                throw std::runtime_error("scene blob: unknown transform kind");
        for (int64_t i = 0; i < np; ++i) {
            const rtb_prim &p = prims()[i];
            if (p.type < 0 || p.type > RTB_PRIM_MEDIUM)
                throw std::runtime_error("scene blob: unknown primitive type");
            if (p.material < 0 || p.material >= nm)
                throw std::runtime_error("scene blob: primitive material out of range");
            if (p.chain < -1 || p.chain >= nc)
                throw std::runtime_error("scene blob: primitive chain out of range");
            if (p.type == RTB_PRIM_MEDIUM) {
                if (p.aux0 < 0 || p.aux1 <= 0 || int64_t(p.aux0) + p.aux1 > np)
                    throw std::runtime_error("scene blob: medium boundary out of range");
                for (int32_t b = p.aux0; b < p.aux0 + p.aux1; ++b)
                    if (prims()[b].type == RTB_PRIM_MEDIUM)
                        throw std::runtime_error("scene blob: nested medium boundary");
            }
        }
        for (int64_t i = 0; i < nm; ++i) {
            const rtb_material &m = materials()[i];
            if (m.type < 0 || m.type >= RTB_MAT_TYPE_COUNT)
                throw std::runtime_error("scene blob: unknown material type");
            for (int k = 0; k < 4; ++k)
                if (m.tex[k] < -1 || m.tex[k] >= nt)
                    throw std::runtime_error("scene blob: material texture out of range");
        }
        for (int64_t i = 0; i < nt; ++i) {
            const rtb_texture &t = textures()[i];
            if (t.type < 0 || t.type > RTB_TEX_NOISE)
                throw std::runtime_error("scene blob: unknown texture type");
            if (t.type == RTB_TEX_CHECKER &&
                (t.even < 0 || t.even >= nt || t.odd < 0 || t.odd >= nt))
                throw std::runtime_error("scene blob: checker child out of range");
            if (t.type == RTB_TEX_IMAGE && (t.image < 0 || t.image >= ni))
                throw std::runtime_error("scene blob: image index out of range");
            if (t.type == RTB_TEX_NOISE && (t.perlin < 0 || t.perlin >= npl))
                throw std::runtime_error("scene blob: perlin index out of range");
        }
        for (int64_t i = 0; i < ni; ++i) {
            const rtb_image &im = images()[i];
            if (im.width < 0 || im.height < 0 || im.offset > n_image_bytes() ||
                uint64_t(im.width) * uint64_t(im.height) > (n_image_bytes() - im.offset) / 3)
                throw std::runtime_error("scene blob: image bytes out of range");
        }
        for (uint64_t i = 0; i < n_lights(); ++i) {
            const rtb_light &l = lights()[i];
            if (l.type < 0 || l.type > RTB_LIGHT_ENV)
                throw std::runtime_error("scene blob: unknown light type");
            if (l.type == RTB_LIGHT_ENV) { // width and height both 0 (the white fallback) or both positive
                if (l.env_width < 0 || l.env_height < 0 || (l.env_width == 0) != (l.env_height == 0))
                    throw std::runtime_error("scene blob: env map size invalid");
                if (l.env_offset > n_env_texels() ||
                    uint64_t(l.env_width) * uint64_t(l.env_height) > (n_env_texels() - l.env_offset) / 3)
                    throw std::runtime_error("scene blob: env texels out of range");
            }
        }
        for (uint64_t i = 0; i < n_gates(); ++i) {
            const rtb_gate &g = gates()[i];
            if (g.prim < 0 || g.prim >= np || prims()[g.prim].type != RTB_PRIM_SPHERE ||
                !(prims()[g.prim].flags & RTB_PRIM_GATED))
                throw std::runtime_error("scene blob: gate does not name a gated sphere");
        }
        for (int64_t i = 0; i < np; ++i)
            if (prims()[i].flags & RTB_PRIM_GATED) {
                int64_t n = 0;
                for (uint64_t k = 0; k < n_gates(); ++k)
                    n += gates()[k].prim == i;
                if (n != 1)
                    throw std::runtime_error("scene blob: gated primitive without exactly one gate");
            }
    }

  private:
    const rtb_blob_section *find(uint32_t id) const {
        for (const auto &s : dir_)
            if (s.id == id)
                return &s;
        return nullptr;
    }
    const uint8_t *base_;
    rtb_blob_header hdr_{};
    std::vector<rtb_blob_section> dir_;
};

} // namespace rtb

#endif // RTB200_BLOB_HPP
