#include "xml_mini.h"

#include <cctype>
#include <cstdio>
#include <cstring>
#include <strings.h>

namespace rtu {

bool ieq(const char *a, const char *b) { return strcasecmp(a, b) == 0; }

const char *XmlElement::attribute(const char *key) const
{
    for (const auto &kv : attrs)
        if (kv.first == key) return kv.second.c_str();
    return nullptr;
}

bool XmlElement::query_double(const char *key, double *out) const
{
    const char *v = attribute(key);
    if (!v) return false;
    double d;
    if (sscanf(v, "%lf", &d) != 1) return false;
    *out = d;
    return true;
}

bool XmlElement::query_int(const char *key, int *out) const
{
    const char *v = attribute(key);
    if (!v) return false;
    int d;
    if (sscanf(v, "%d", &d) != 1) return false;
    *out = d;
    return true;
}

const XmlElement *XmlElement::first_child(const char *name_ci) const
{
    for (const auto &c : children)
        if (ieq(c->name.c_str(), name_ci)) return c.get();
    return nullptr;
}

bool xml_read_file(const char *path, std::string *out)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return false;
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) s.append(buf, n);
    fclose(fp);
    *out = std::move(s);
    return true;
}

namespace {

struct Parser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string &t) : s(t) {}

    bool starts(const char *lit) const { return s.compare(i, strlen(lit), lit) == 0; }
    void skip_ws()
    {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
    }
    static bool name_char(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }

    std::string decode(const std::string &v)
    {
        if (v.find('&') == std::string::npos) return v;
        std::string o;
        for (size_t k = 0; k < v.size(); k++) {
            if (v[k] == '&') {
                if (v.compare(k, 5, "&amp;") == 0) { o += '&'; k += 4; continue; }
                if (v.compare(k, 4, "&lt;") == 0) { o += '<'; k += 3; continue; }
                if (v.compare(k, 4, "&gt;") == 0) { o += '>'; k += 3; continue; }
                if (v.compare(k, 6, "&quot;") == 0) { o += '"'; k += 5; continue; }
                if (v.compare(k, 6, "&apos;") == 0) { o += '\''; k += 5; continue; }
            }
            o += v[k];
        }
        return o;
    }

    // skips text, comments, declarations, doctype up to the next '<' that opens/closes an element
    bool skip_misc()
    {
        for (;;) {
            while (i < s.size() && s[i] != '<') i++;
            if (i >= s.size()) return true;
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) { err = "unterminated declaration"; return false; }
                i = e + 2;
            } else if (starts("<!")) {
                size_t e = s.find('>', i + 2);
                if (e == std::string::npos) { err = "unterminated <! block"; return false; }
                i = e + 1;
            } else {
                return true;
            }
        }
    }

    bool parse_children(XmlElement *parent, const std::string *closing)
    {
        for (;;) {
            if (!skip_misc()) return false;
            if (i >= s.size()) {
                if (closing) { err = "missing </" + *closing + ">"; return false; }
                return true;
            }
            if (starts("</")) {
                i += 2;
                size_t b = i;
                while (i < s.size() && name_char(s[i])) i++;
                std::string nm = s.substr(b, i - b);
                skip_ws();
                if (i >= s.size() || s[i] != '>') { err = "malformed closing tag"; return false; }
                i++;
                if (!closing || nm != *closing) { err = "unexpected </" + nm + ">"; return false; }
                return true;
            }
            // opening tag
            i++;
            size_t b = i;
            while (i < s.size() && name_char(s[i])) i++;
            if (i == b) { err = "malformed tag"; return false; }
            std::unique_ptr<XmlElement> el(new XmlElement);
            el->name = s.substr(b, i - b);
            bool selfclose = false;
            for (;;) {
                skip_ws();
                if (i >= s.size()) { err = "unterminated tag <" + el->name; return false; }
                if (s[i] == '>') { i++; break; }
                if (s[i] == '/' && i + 1 < s.size() && s[i + 1] == '>') { i += 2; selfclose = true; break; }
                size_t ab = i;
                while (i < s.size() && name_char(s[i])) i++;
                if (i == ab) { err = "malformed attribute in <" + el->name + ">"; return false; }
                std::string key = s.substr(ab, i - ab);
                skip_ws();
                std::string val;
                if (i < s.size() && s[i] == '=') {
                    i++;
                    skip_ws();
                    if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
                        char q = s[i++];
                        size_t vb = i;
                        while (i < s.size() && s[i] != q) i++;
                        if (i >= s.size()) { err = "unterminated attribute value"; return false; }
                        val = s.substr(vb, i - vb);
                        i++;
                    } else { // unquoted value (TinyXML tolerates it)
                        size_t vb = i;
                        while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') i++;
                        val = s.substr(vb, i - vb);
                    }
                }
                el->attrs.emplace_back(key, decode(val));
            }
            XmlElement *raw = el.get();
            parent->children.push_back(std::move(el));
            if (!selfclose) {
                std::string nm = raw->name;
                if (!parse_children(raw, &nm)) return false;
            }
        }
    }
};

} // namespace

std::unique_ptr<XmlElement> xml_parse(const std::string &text, std::string *err)
{
    Parser p(text);
    std::unique_ptr<XmlElement> root(new XmlElement);
    if (!p.parse_children(root.get(), nullptr)) {
        if (err) *err = p.err;
        return nullptr;
    }
    return root;
}

} // namespace rtu
