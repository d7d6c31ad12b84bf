// Scene front-end: the reference's XML scene grammar -> rtu_scene_desc.
//
// Follows LoadScene and its helpers (ExternalLibrary/xmlload.cpp:64-553) element by element:
// same defaults, same attribute names, same order of transform composition, same camera
// fix-up, same "first material with that name wins" binding.  The result is the flattened,
// pre-order node table of include/rtu.h instead of the reference's pointer graph.
#include <cstdio>
#include <cstring>
#include <map>

#include "hmath.h"
#include "host_scene.h"
#include "xml_mini.h"

namespace rtu {

namespace {

thread_local std::string g_last_error;

struct Loader {
    rtu_host_scene *hs;
    uint32_t flags = 0; // RTU_LOAD_*
    std::string root; // asset root for relative paths
    struct Pending { int node; std::string mtl; };
    std::vector<Pending> node_mtl;              // nodeMtlList (xmlload.cpp:54-60)
    std::map<std::string, int> mesh_by_name;    // objList.Find (xmlload.cpp:201)
    std::map<std::string, int> tex_by_name;     // textureList.Find (xmlload.cpp:537)
    std::vector<std::string> warnings;

    std::string resolve(const char *name) const
    {
        if (name[0] == '/' || root.empty()) return name;
        return root + "/" + name;
    }

    // ---- ReadFloat / ReadVector / ReadColor (xmlload.cpp:452-495)
    static void read_float(const XmlElement *e, float &f, const char *name = "value")
    {
        double d = (double)f;
        e->query_double(name, &d);
        f = (float)d;
    }
    static void read_vector(const XmlElement *e, V3 &v)
    {
        double x = v.x, y = v.y, z = v.z;
        e->query_double("x", &x);
        e->query_double("y", &y);
        e->query_double("z", &z);
        v.x = (float)x; v.y = (float)y; v.z = (float)z;
        float f = 1;
        read_float(e, f);
        v = v * f;
    }
    static void read_color(const XmlElement *e, float c[3])
    {
        double r = c[0], g = c[1], b = c[2];
        e->query_double("r", &r);
        e->query_double("g", &g);
        e->query_double("b", &b);
        c[0] = (float)r; c[1] = (float)g; c[2] = (float)b;
        float f = 1;
        read_float(e, f);
        c[0] *= f; c[1] *= f; c[2] *= f;
    }

    // ---- LoadTransform (xmlload.cpp:264-290)
    static void load_transform(Xform &t, const XmlElement *e)
    {
        for (const auto &c : e->children) {
            const char *tag = c->name.c_str();
            if (ieq(tag, "scale")) {
                V3 s(1, 1, 1);
                read_vector(c.get(), s);
                t.scale(s.x, s.y, s.z);
            } else if (ieq(tag, "rotate")) {
                V3 s(0, 0, 0);
                read_vector(c.get(), s);
                s = normalized(s);
                float a = 0; // the reference leaves this uninitialised when "angle" is missing
                read_float(c.get(), a, "angle");
                t.rotate(s, a);
            } else if (ieq(tag, "translate")) {
                V3 p(0, 0, 0);
                read_vector(c.get(), p);
                t.translate(p);
            }
        }
    }

    // ---- ReadTexture(const char*) (xmlload.cpp:534-553): -1 when the file fails to load
    int file_texture(const char *name)
    {
        auto it = tex_by_name.find(name);
        if (it != tex_by_name.end()) return it->second;
        std::unique_ptr<HostTexture> t(new HostTexture);
        t->name = name;
        t->kind = RTU_TEX_FILE;
        size_t len = strlen(name);
        bool ok = false;
        std::string err;
        if (len >= 3) {
            char ext[4] = {(char)tolower(name[len - 3]), (char)tolower(name[len - 2]), (char)tolower(name[len - 1]), 0};
            if (!strcmp(ext, "png")) ok = decode_png_rgb8(resolve(name).c_str(), &t->rgb8, &t->width, &t->height, &err);
            else if (!strcmp(ext, "ppm")) ok = decode_ppm_rgb8(resolve(name).c_str(), &t->rgb8, &t->width, &t->height, &err);
            else err = "unsupported texture format"; // TextureFile::Load knows png and ppm (texture.cpp:71-88)
        }
        if (!ok) {
            warnings.push_back(std::string("texture \"") + name + "\" not loaded: " + err);
            return -1;
        }
        hs->textures.push_back(std::move(t));
        int id = (int)hs->textures.size() - 1;
        tex_by_name[name] = id;
        return id;
    }

    // ---- ReadTexture(TiXmlElement*) (xmlload.cpp:499-530): index into texmaps, -1 = no map
    int read_texture(const XmlElement *e)
    {
        const char *tex_name = e->attribute("texture");
        if (!tex_name) return -1;
        int tex = -1;
        if (ieq(tex_name, "checkerboard")) {
            std::unique_ptr<HostTexture> t(new HostTexture);
            t->name = tex_name;
            t->kind = RTU_TEX_CHECKER;
            for (const auto &c : e->children) {
                if (ieq(c->name.c_str(), "color1")) { float col[3] = {0, 0, 0}; read_color(c.get(), col); memcpy(t->color1, col, sizeof col); }
                else if (ieq(c->name.c_str(), "color2")) { float col[3] = {0, 0, 0}; read_color(c.get(), col); memcpy(t->color2, col, sizeof col); }
            }
            hs->textures.push_back(std::move(t));
            tex = (int)hs->textures.size() - 1;
        } else {
            tex = file_texture(tex_name);
        }
        Xform x;
        load_transform(x, e);
        rtu_texmap m;
        memset(&m, 0, sizeof m);
        m.kind = RTU_TEX_NULL;
        memcpy(m.itm, x.itm.d, sizeof m.itm);
        m.pos[0] = x.pos.x; m.pos[1] = x.pos.y; m.pos[2] = x.pos.z;
        if (tex >= 0) {
            const HostTexture &t = *hs->textures[tex];
            m.kind = t.kind;
            memcpy(m.color1, t.color1, sizeof m.color1);
            memcpy(m.color2, t.color2, sizeof m.color2);
            m.rgb8 = t.rgb8.empty() ? nullptr : t.rgb8.data();
            m.width = t.width;
            m.height = t.height;
        }
        hs->texmaps.push_back(m);
        return (int)hs->texmaps.size() - 1;
    }

    // new TextureMap(ReadTexture(filename)) (xmlload.cpp:221-225): identity transform; a file that fails to load leaves a
    // map without texture, which samples black (scene.h:382)
    int file_texmap(const char *file)
    {
        int tex = file_texture(file);
        Xform x;
        rtu_texmap m;
        memset(&m, 0, sizeof m);
        m.kind = RTU_TEX_NULL;
        memcpy(m.itm, x.itm.d, sizeof m.itm);
        if (tex >= 0) {
            const HostTexture &t = *hs->textures[tex];
            m.kind = t.kind;
            memcpy(m.color1, t.color1, sizeof m.color1);
            memcpy(m.color2, t.color2, sizeof m.color2);
            m.rgb8 = t.rgb8.empty() ? nullptr : t.rgb8.data();
            m.width = t.width;
            m.height = t.height;
        }
        hs->texmaps.push_back(m);
        return (int)hs->texmaps.size() - 1;
    }

    // The MultiMtl xmlload.cpp:208-241 generates for an OBJ that brought its own materials, as the only sub-material that
    // can ever shade: hInfo.mtlID is never set by IntersectTriangle, so MultiMtl::Shade (materials.h:66) always takes
    // mtls[0].  The textures of every sub-material are still read, in order, like the reference does (textureList order).
    void obj_multi_material(const HostMesh &hm, const char *name, int node)
    {
        for (const std::string &n : hs->material_names) if (n == name) return; // materials.Find(name) != NULL
        rtu_material first;
        bool have = false;
        for (const ObjMtl &mtl : hm.mtls) {
            rtu_material m; // MtlBlinn() defaults, materials.h:23-25
            m.diffuse = texcolor(mtl.Kd[0], mtl.Kd[1], mtl.Kd[2]);
            m.specular = texcolor(mtl.Ks[0], mtl.Ks[1], mtl.Ks[2]);
            m.reflection = texcolor(0, 0, 0);
            m.refraction = texcolor(0, 0, 0);
            m.emission = texcolor(0, 0, 0);
            m.glossiness = mtl.Ns;
            m.absorption[0] = m.absorption[1] = m.absorption[2] = 0;
            m.ior = mtl.Ni;
            m.reflection_glossiness = m.refraction_glossiness = 0;
            if (mtl.has_map_Kd) m.diffuse.texmap = file_texmap(mtl.map_Kd.c_str());
            if (mtl.has_map_Ks) m.diffuse.texmap = file_texmap(mtl.map_Ks.c_str()); // sic: SetDiffuseTexture (xmlload.cpp:222, SURVEY A-19)
            if (mtl.illum > 2 && mtl.illum <= 7) {
                m.reflection = texcolor(mtl.Ks[0], mtl.Ks[1], mtl.Ks[2]);
                if (mtl.has_map_Ks) m.reflection.texmap = file_texmap(mtl.map_Ks.c_str());
                if (mtl.illum >= 6) m.refraction = texcolor(1 - mtl.Tf[0], 1 - mtl.Tf[1], 1 - mtl.Tf[2]);
            }
            if (!have) { first = m; have = true; }
        }
        if (!have) return;
        hs->materials.push_back(first);
        hs->material_names.push_back(name);
        node_mtl.push_back({node, name});
    }

    static rtu_texcolor texcolor(float r, float g, float b)
    {
        rtu_texcolor c;
        c.color[0] = r; c.color[1] = g; c.color[2] = b;
        c.texmap = -1;
        return c;
    }

    // ---- LoadMaterial (xmlload.cpp:294-370)
    void load_material(const XmlElement *e)
    {
        const char *name = e->attribute("name");
        const char *type = e->attribute("type");
        if (!type || !ieq(type, "blinn")) return;
        rtu_material m; // MtlBlinn() defaults, materials.h:23-25
        m.diffuse = texcolor(0.5f, 0.5f, 0.5f);
        m.specular = texcolor(0.7f, 0.7f, 0.7f);
        m.reflection = texcolor(0, 0, 0);
        m.refraction = texcolor(0, 0, 0);
        m.emission = texcolor(0, 0, 0);
        m.glossiness = 20.0f;
        m.absorption[0] = m.absorption[1] = m.absorption[2] = 0;
        m.ior = 1;
        m.reflection_glossiness = m.refraction_glossiness = 0;
        for (const auto &c : e->children) {
            float col[3] = {1, 1, 1};
            float f = 1;
            const char *tag = c->name.c_str();
            if (ieq(tag, "diffuse")) {
                read_color(c.get(), col);
                memcpy(m.diffuse.color, col, sizeof col);
                m.diffuse.texmap = read_texture(c.get());
            } else if (ieq(tag, "specular")) {
                read_color(c.get(), col);
                memcpy(m.specular.color, col, sizeof col);
                m.specular.texmap = read_texture(c.get());
            } else if (ieq(tag, "glossiness")) {
                read_float(c.get(), f);
                m.glossiness = f;
            } else if (ieq(tag, "emission")) {
                read_color(c.get(), col);
                memcpy(m.emission.color, col, sizeof col);
                m.emission.texmap = read_texture(c.get());
            } else if (ieq(tag, "reflection")) {
                read_color(c.get(), col);
                memcpy(m.reflection.color, col, sizeof col);
                m.reflection.texmap = read_texture(c.get());
                f = 0;
                read_float(c.get(), f, "glossiness");
                m.reflection_glossiness = f;
            } else if (ieq(tag, "refraction")) {
                read_color(c.get(), col);
                memcpy(m.refraction.color, col, sizeof col);
                read_float(c.get(), f, "index");
                m.ior = f;
                m.refraction.texmap = read_texture(c.get());
                f = 0;
                read_float(c.get(), f, "glossiness");
                m.refraction_glossiness = f;
            } else if (ieq(tag, "absorption")) {
                read_color(c.get(), col);
                memcpy(m.absorption, col, sizeof col);
            }
        }
        hs->materials.push_back(m);
        hs->material_names.push_back(name ? name : "");
    }

    // ---- LoadLight (xmlload.cpp:374-448)
    void load_light(const XmlElement *e)
    {
        const char *type = e->attribute("type");
        if (!type) return;
        rtu_light l;
        memset(&l, 0, sizeof l);
        if (ieq(type, "ambient")) {
            l.kind = RTU_LIGHT_AMBIENT;
            for (const auto &c : e->children)
                if (ieq(c->name.c_str(), "intensity")) { float col[3] = {1, 1, 1}; read_color(c.get(), col); memcpy(l.intensity, col, sizeof col); }
        } else if (ieq(type, "direct")) {
            l.kind = RTU_LIGHT_DIRECT;
            l.v[2] = 1; // DirectLight(): direction (0,0,1), lights.h:47
            for (const auto &c : e->children) {
                if (ieq(c->name.c_str(), "intensity")) { float col[3] = {1, 1, 1}; read_color(c.get(), col); memcpy(l.intensity, col, sizeof col); }
                else if (ieq(c->name.c_str(), "direction")) {
                    V3 v(1, 1, 1);
                    read_vector(c.get(), v);
                    v = normalized(v); // SetDirection, lights.h:53
                    l.v[0] = v.x; l.v[1] = v.y; l.v[2] = v.z;
                }
            }
        } else if (ieq(type, "point")) {
            l.kind = RTU_LIGHT_POINT;
            for (const auto &c : e->children) {
                if (ieq(c->name.c_str(), "intensity")) { float col[3] = {1, 1, 1}; read_color(c.get(), col); memcpy(l.intensity, col, sizeof col); }
                else if (ieq(c->name.c_str(), "position")) { V3 v(0, 0, 0); read_vector(c.get(), v); l.v[0] = v.x; l.v[1] = v.y; l.v[2] = v.z; }
                else if (ieq(c->name.c_str(), "size")) { float f = 0; read_float(c.get(), f); l.size = f; }
            }
        } else {
            return;
        }
        hs->lights.push_back(l);
    }

    // ---- LoadNode (xmlload.cpp:167-260).  Pre-order numbering: the node gets its index
    // before its children, which is the order Trace() visits them.
    void load_node(int parent, const XmlElement *e)
    {
        int me = (int)hs->nodes.size();
        rtu_node n;
        memset(&n, 0, sizeof n);
        n.parent = parent;
        n.kind = RTU_OBJ_NONE;
        n.mesh = -1;
        n.material = -1;
        hs->nodes.push_back(n);
        const char *name = e->attribute("name");
        hs->node_names.push_back(name ? name : "");
        const char *mtl = e->attribute("material");
        if (mtl) node_mtl.push_back({me, mtl});
        const char *type = e->attribute("type");
        if (type) {
            if (ieq(type, "sphere")) hs->nodes[me].kind = RTU_OBJ_SPHERE;
            else if (ieq(type, "plane")) hs->nodes[me].kind = RTU_OBJ_PLANE;
            else if (ieq(type, "obj") && name) {
                auto it = mesh_by_name.find(name);
                int mesh = -1;
                if (it != mesh_by_name.end()) {
                    mesh = it->second;
                } else {
                    std::unique_ptr<HostMesh> hm(new HostMesh);
                    std::string err;
                    if (load_obj_mesh(resolve(name).c_str(), hm.get(), &err, mtl == nullptr, !(flags & RTU_LOAD_DEVICE_BVH))) {
                        hm->name = name;
                        if (!hm->mtls.empty()) obj_multi_material(*hm, name, me); // only on the first load of this OBJ (xmlload.cpp:201-241)
                        hs->meshes.push_back(std::move(hm));
                        mesh = (int)hs->meshes.size() - 1;
                        mesh_by_name[name] = mesh;
                    } else {
                        warnings.push_back(err); // xmlload.cpp:205: the node simply has no object
                    }
                }
                if (mesh >= 0) { hs->nodes[me].kind = RTU_OBJ_MESH; hs->nodes[me].mesh = mesh; }
            }
        }
        for (const auto &c : e->children)
            if (ieq(c->name.c_str(), "object")) load_node(me, c.get());
        Xform x;
        load_transform(x, e);
        rtu_node &nn = hs->nodes[me];
        memcpy(nn.tm, x.tm.d, sizeof nn.tm);
        memcpy(nn.itm, x.itm.d, sizeof nn.itm);
        nn.pos[0] = x.pos.x; nn.pos[1] = x.pos.y; nn.pos[2] = x.pos.z;
    }

    bool load(const char *xml_path)
    {
        std::string text, err;
        if (!xml_read_file(xml_path, &text)) { set_error(std::string("Failed to load the file \"") + xml_path + "\""); return false; }
        std::unique_ptr<XmlElement> doc = xml_parse(text, &err);
        if (!doc) { set_error(std::string(xml_path) + ": XML syntax error: " + err); return false; }
        const XmlElement *xml = nullptr, *scene = nullptr, *cam = nullptr;
        for (const auto &c : doc->children) if (c->name == "xml") { xml = c.get(); break; }
        if (!xml) { set_error("No \"xml\" tag found."); return false; }
        for (const auto &c : xml->children) if (c->name == "scene") { scene = c.get(); break; }
        if (!scene) { set_error("No \"scene\" tag found."); return false; }
        for (const auto &c : xml->children) if (c->name == "camera") { cam = c.get(); break; }
        if (!cam) { set_error("No \"camera\" tag found."); return false; }

        // rootNode.Init(): identity transform, no object (xmlload.cpp:91, scene.h:449)
        rtu_node rootn;
        memset(&rootn, 0, sizeof rootn);
        Xform id;
        memcpy(rootn.tm, id.tm.d, sizeof rootn.tm);
        memcpy(rootn.itm, id.itm.d, sizeof rootn.itm);
        rootn.parent = -1;
        rootn.kind = RTU_OBJ_NONE;
        rootn.mesh = -1;
        rootn.material = -1;
        hs->nodes.push_back(rootn);
        hs->node_names.push_back("");
        hs->desc.background = texcolor(0, 0, 0);  // TexturedColor(): colour 0, no map (scene.h:411)
        hs->desc.environment = texcolor(0, 0, 0);

        for (const auto &c : scene->children) { // LoadScene(TiXmlElement*) xmlload.cpp:139-163
            const char *tag = c->name.c_str();
            if (ieq(tag, "background")) {
                float col[3] = {1, 1, 1};
                read_color(c.get(), col);
                memcpy(hs->desc.background.color, col, sizeof col);
                hs->desc.background.texmap = read_texture(c.get());
            } else if (ieq(tag, "environment")) {
                float col[3] = {1, 1, 1};
                read_color(c.get(), col);
                memcpy(hs->desc.environment.color, col, sizeof col);
                hs->desc.environment.texmap = read_texture(c.get());
            } else if (ieq(tag, "object")) {
                load_node(0, c.get());
            } else if (ieq(tag, "material")) {
                load_material(c.get());
            } else if (ieq(tag, "light")) {
                load_light(c.get());
            }
        }
        // material binding (xmlload.cpp:100-106): first material with that name; later list
        // entries for the same node override earlier ones
        for (const auto &p : node_mtl)
            for (size_t k = 0; k < hs->material_names.size(); k++)
                if (hs->material_names[k] == p.mtl) { hs->nodes[p.node].material = (int)k; break; }

        // camera (xmlload.cpp:108-126)
        rtu_camera &cm = hs->desc.camera;
        V3 pos(0, 0, 0), dir(0, 0, -1), up(0, 1, 0);
        cm.fov = 40; cm.focaldist = 1; cm.dof = 0; cm.width = 200; cm.height = 150;
        dir = dir + pos;
        for (const auto &c : cam->children) {
            const char *tag = c->name.c_str();
            if (ieq(tag, "position")) read_vector(c.get(), pos);
            else if (ieq(tag, "target")) read_vector(c.get(), dir);
            else if (ieq(tag, "up")) read_vector(c.get(), up);
            else if (ieq(tag, "fov")) read_float(c.get(), cm.fov);
            else if (ieq(tag, "focaldist")) read_float(c.get(), cm.focaldist);
            else if (ieq(tag, "dof")) read_float(c.get(), cm.dof);
            else if (ieq(tag, "width")) c->query_int("value", &cm.width);
            else if (ieq(tag, "height")) c->query_int("value", &cm.height);
        }
        dir = dir - pos;
        dir = normalized(dir);
        V3 x = cross(dir, up);
        up = normalized(cross(x, dir));
        cm.pos[0] = pos.x; cm.pos[1] = pos.y; cm.pos[2] = pos.z;
        cm.dir[0] = dir.x; cm.dir[1] = dir.y; cm.dir[2] = dir.z;
        cm.up[0] = up.x; cm.up[1] = up.y; cm.up[2] = up.z;
        return true;
    }
};

} // namespace

void set_error(const std::string &msg) { g_last_error = msg; }
const std::string &last_error() { return g_last_error; }

} // namespace rtu

void rtu_host_scene::finalize()
{
    mesh_descs.clear();
    for (const auto &m : meshes) {
        rtu_mesh d;
        memset(&d, 0, sizeof d);
        d.v = m->v.data();   d.nv = (uint32_t)(m->v.size() / 3);
        d.vn = m->vn.data(); d.nvn = (uint32_t)(m->vn.size() / 3);
        d.vt = m->vt.empty() ? nullptr : m->vt.data(); d.nvt = (uint32_t)(m->vt.size() / 3);
        d.f = m->f.data();
        d.fn = m->fn.empty() ? nullptr : m->fn.data();
        d.ft = m->ft.empty() ? nullptr : m->ft.data();
        d.nf = m->nf();
        if (m->device_bvh) d.flags |= RTU_MESH_DEVICE_BVH;
        d.bvh_boxes = m->bvh_boxes.empty() ? nullptr : m->bvh_boxes.data();
        d.bvh_data = m->bvh_data.empty() ? nullptr : m->bvh_data.data();
        d.bvh_nodes = (uint32_t)m->bvh_data.size();
        d.bvh_elements = m->bvh_elements.empty() ? nullptr : m->bvh_elements.data();
        if (!m->occ.slots.empty()) {
            d.occ_nodes = m->occ.nodes.empty() ? nullptr : m->occ.nodes.data();
            d.occ_n_nodes = (uint32_t)(m->occ.nodes.size() / 32);
            d.occ_root = m->occ.root;
            d.occ_slots = m->occ.slots.data();
        }
        memcpy(d.bound_min, m->bound_min, sizeof d.bound_min);
        memcpy(d.bound_max, m->bound_max, sizeof d.bound_max);
        mesh_descs.push_back(d);
    }
    desc.nodes = nodes.data();         desc.n_nodes = (int32_t)nodes.size();
    desc.meshes = mesh_descs.data();   desc.n_meshes = (int32_t)mesh_descs.size();
    desc.materials = materials.data(); desc.n_materials = (int32_t)materials.size();
    desc.lights = lights.data();       desc.n_lights = (int32_t)lights.size();
    desc.texmaps = texmaps.data();     desc.n_texmaps = (int32_t)texmaps.size();
    // light masks (light_mask.cpp), once per load instead of once per rtu_scene_upload
    desc.light_masks = nullptr;
    desc.n_light_masks = 0;
    rtu::collect_light_masks(desc, &light_masks, &light_mask_data, 1ull << 28);
    desc.light_masks = light_masks.empty() ? nullptr : light_masks.data();
    desc.n_light_masks = (int32_t)light_masks.size();
}

extern "C" {

const char *rtu_last_error(void)
{
    static thread_local std::string copy;
    copy = rtu::last_error();
    return copy.c_str();
}

int rtu_version(void) { return 1; }

int rtu_host_load_xml(const char *xml_path, const char *asset_root, rtu_host_scene **out)
{
    return rtu_host_load_xml_ex(xml_path, asset_root, 0u, out);
}

int rtu_host_load_xml_ex(const char *xml_path, const char *asset_root, uint32_t flags, rtu_host_scene **out)
{
    if (!xml_path || !out) { rtu::set_error("rtu_host_load_xml: null argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_host_load_xml", [&]() -> int {
    std::unique_ptr<rtu_host_scene> hs(new rtu_host_scene);
    memset(&hs->desc, 0, sizeof hs->desc);
    rtu::Loader L;
    L.hs = hs.get();
    L.flags = flags;
    L.root = asset_root ? asset_root : "";
    if (!L.load(xml_path)) return RTU_ERR_IO;
    // texmaps captured rgb8 pointers while the texture list was still growing; the pixel
    // vectors live in unique_ptr-owned HostTextures, so those pointers stay valid.
    hs->finalize();
    if (!L.warnings.empty()) {
        std::string w;
        for (const auto &s : L.warnings) w += s + "; ";
        rtu::set_error("warnings: " + w);
    } else {
        rtu::set_error("");
    }
    *out = hs.release();
    return RTU_OK;
    });
}

const rtu_scene_desc *rtu_host_scene_desc(const rtu_host_scene *hs) { return hs ? &hs->desc : nullptr; }

void rtu_host_scene_destroy(rtu_host_scene *hs) { delete hs; }

int rtu_host_build_bvh(const float *v, uint32_t nv, const uint32_t *f, uint32_t nf, uint32_t max_per_leaf,
                       float *boxes, uint32_t *data, uint32_t *elements, uint32_t *n_nodes)
{
    if (!v || !f || !boxes || !data || !elements || !n_nodes) { rtu::set_error("rtu_host_build_bvh: null argument"); return RTU_ERR_INVALID; }
    for (uint32_t i = 0; i < nf * 3; i++)
        if (f[i] >= nv) { rtu::set_error("rtu_host_build_bvh: face index out of range"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_host_build_bvh", [&]() -> int {
    std::vector<float> b;
    std::vector<uint32_t> d, e;
    rtu::build_bvh(v, f, nf, max_per_leaf, &b, &d, &e);
    memcpy(boxes, b.data(), b.size() * sizeof(float));
    memcpy(data, d.data(), d.size() * sizeof(uint32_t));
    memcpy(elements, e.data(), e.size() * sizeof(uint32_t));
    *n_nodes = (uint32_t)d.size();
    return RTU_OK;
    });
}

int rtu_host_build_occlusion_bvh(const float *v, uint32_t nv, const uint32_t *f, uint32_t nf, const uint32_t *bvh_elements,
                                 float *nodes, uint32_t *n_nodes, uint32_t *root, uint32_t *slots)
{
    if (!v || !f || !bvh_elements || !nodes || !n_nodes || !root || !slots) { rtu::set_error("rtu_host_build_occlusion_bvh: null argument"); return RTU_ERR_INVALID; }
    for (uint32_t i = 0; i < nf * 3; i++)
        if (f[i] >= nv) { rtu::set_error("rtu_host_build_occlusion_bvh: face index out of range"); return RTU_ERR_INVALID; }
    for (uint32_t i = 0; i < nf; i++)
        if (bvh_elements[i] >= nf) { rtu::set_error("rtu_host_build_occlusion_bvh: element out of range"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_host_build_occlusion_bvh", [&]() -> int {
        rtu::OccBvh o;
        rtu::build_occlusion_bvh(v, f, bvh_elements, nf, &o);
        memcpy(nodes, o.nodes.data(), o.nodes.size() * sizeof(float));
        memcpy(slots, o.slots.data(), o.slots.size() * sizeof(uint32_t));
        *n_nodes = (uint32_t)(o.nodes.size() / 32);
        *root = o.root;
        return RTU_OK;
    });
}

int rtu_host_build_light_mask(const rtu_scene_desc *d, int32_t node, int32_t light, float *rec, uint32_t *bits)
{
    if (!d || !rec || !bits) { rtu::set_error("rtu_host_build_light_mask: null argument"); return RTU_ERR_INVALID; }
    if (node < 0 || node >= d->n_nodes || light < -1 || light >= d->n_lights || d->nodes[node].kind != RTU_OBJ_MESH ||
        d->nodes[node].mesh < 0 || d->nodes[node].mesh >= d->n_meshes) {
        rtu::set_error("rtu_host_build_light_mask: not a mesh node / light of this scene");
        return RTU_ERR_INVALID;
    }
    return rtu::guarded("rtu_host_build_light_mask", [&]() -> int {
        std::vector<const rtu_node *> chain;
        for (int a = node; a >= 0 && chain.size() < 64; a = d->nodes[a].parent) chain.insert(chain.begin(), &d->nodes[a]);
        std::vector<uint32_t> b;
        rtu_light eye; // light -1: the camera's mask
        memset(&eye, 0, sizeof eye);
        eye.kind = RTU_LIGHT_POINT;
        memcpy(eye.v, d->camera.pos, sizeof eye.v);
        if ((light < 0 && d->camera.dof > 0.f) ||
            !rtu::build_light_mask(chain.data(), (int)chain.size(), d->meshes[d->nodes[node].mesh], light < 0 ? eye : d->lights[light], rec, &b, light < 0)) {
            rtu::set_error("rtu_host_build_light_mask: no mask for this mesh and light");
            return RTU_ERR_UNSUPPORTED;
        }
        memcpy(bits, b.data(), b.size() * sizeof(uint32_t));
        return RTU_OK;
    });
}

int rtu_write_png(const char *path, const uint8_t *pixels, int32_t width, int32_t height, int32_t channels)
{
    std::string err;
    if (!path || !pixels) { rtu::set_error("rtu_write_png: null argument"); return RTU_ERR_INVALID; }
    return rtu::guarded("rtu_write_png", [&]() -> int {
        if (!rtu::encode_png(path, pixels, width, height, channels, &err)) { rtu::set_error(err); return RTU_ERR_IO; }
        return RTU_OK;
    });
}

} // extern "C"
