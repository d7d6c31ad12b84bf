// Minimal XML DOM, just enough for the reference's scene grammar (xmlload.cpp): elements,
// attributes (single or double quoted), comments, declarations, self-closing tags.  Text
// content is ignored (the scene grammar never uses it).  Query semantics follow what
// xmlload.cpp relies on from TinyXML 2.6.2: Attribute() returns NULL when absent, the
// numeric queries leave the output untouched when the attribute is absent or not a number,
// element names are compared case-insensitively by the caller (COMPARE, xmlload.cpp:33-37).
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace rtu {

struct XmlElement {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlElement>> children;

    const char *attribute(const char *key) const;
    bool query_double(const char *key, double *out) const; // sscanf("%lf") like TinyXML
    bool query_int(const char *key, int *out) const;
    const XmlElement *first_child(const char *name_ci) const; // case-insensitive
};

// Parses a whole document; returns a synthetic root whose children are the top-level
// elements.  On a syntax error returns nullptr and fills err.
std::unique_ptr<XmlElement> xml_parse(const std::string &text, std::string *err);
bool xml_read_file(const char *path, std::string *out);
bool ieq(const char *a, const char *b);

} // namespace rtu
