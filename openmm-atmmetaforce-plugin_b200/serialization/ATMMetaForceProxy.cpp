#include "ATMMetaForceProxy.h"

#include <cctype>
#include <memory>

#include "ATMMetaForce.h"

using namespace ATMMetaForcePlugin;
using OpenMM::SerializationNode;

ATMMetaForceProxy::ATMMetaForceProxy() : OpenMM::SerializationProxy("ATMMetaForce") {}

void ATMMetaForceProxy::serialize(const void *object, SerializationNode &node) const {
    const ATMMetaForce &f = *static_cast<const ATMMetaForce *>(object);
    node.setIntProperty("version", 0);
    node.setIntProperty("forceGroup", f.getForceGroup());
    node.setStringProperty("name", f.getName());
    double p[9];
    f.getDefaultParameters(p);
    static const char *keys[9] = {"lambda1", "lambda2", "alpha", "u0", "w0", "uMax", "ubCore", "aCore", "direction"};
    for (int k = 0; k < 9; k++) node.setDoubleProperty(keys[k], p[k]);
    SerializationNode &groups = node.createChildNode("VariableForceGroups");
    for (int g : f.getVariableForceGroups()) groups.createChildNode("Parameter").setIntProperty("group", g);
    SerializationNode &parts = node.createChildNode("Particles");
    for (int i = 0; i < f.getNumParticles(); i++) {
        int particle;
        double dx, dy, dz;
        f.getParticleParameters(i, particle, dx, dy, dz);
        parts.createChildNode("Particle").setIntProperty("particle", particle).setDoubleProperty("dx", dx)
            .setDoubleProperty("dy", dy).setDoubleProperty("dz", dz);
    }
}

void *ATMMetaForceProxy::deserialize(const SerializationNode &node) const {
    if (node.getIntProperty("version") != 0) throw OpenMM::OpenMMException("Unsupported version");
    std::vector<int> groups;
    for (const SerializationNode &g : node.getChildNode("VariableForceGroups").getChildren())
        groups.push_back(g.getIntProperty("group"));
    std::unique_ptr<ATMMetaForce> f(new ATMMetaForce(
        node.getDoubleProperty("lambda1"), node.getDoubleProperty("lambda2"), node.getDoubleProperty("alpha"),
        node.getDoubleProperty("u0"), node.getDoubleProperty("w0"), node.getDoubleProperty("uMax"),
        node.getDoubleProperty("ubCore"), node.getDoubleProperty("aCore"), node.getDoubleProperty("direction"), groups));
    f->setForceGroup(node.getIntProperty("forceGroup", 0));
    f->setName(node.getStringProperty("name", f->getName()));
    for (const SerializationNode &p : node.getChildNode("Particles").getChildren())
        f->addParticle(p.getIntProperty("particle"), p.getDoubleProperty("dx"), p.getDoubleProperty("dy"),
                       p.getDoubleProperty("dz"));
    return f.release();
}

void ATMMetaForcePlugin::registerATMMetaForceSerializationProxies() {
    static ATMMetaForceProxy proxy;
    OpenMM::SerializationProxy::registerProxy(&proxy);
}

namespace {
struct Registrar {
    Registrar() { registerATMMetaForceSerializationProxies(); }
} registrar;  // static-constructor registration, as the reference does
}  // namespace

#ifndef ATM_HAVE_OPENMM
// ---------------------------------------------------------------------------------------------------------------
// Minimal writer / reader for OpenMM's XML dialect (only used when OpenMM's own XmlSerializer is not available).
// ---------------------------------------------------------------------------------------------------------------
namespace OpenMM {

static std::string escape(const std::string &s) {
    std::string o;
    for (char c : s) {
        switch (c) {
            case '&': o += "&amp;"; break;
            case '<': o += "&lt;"; break;
            case '>': o += "&gt;"; break;
            case '"': o += "&quot;"; break;
            default: o += c;
        }
    }
    return o;
}

static std::string unescape(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '&') {
            if (s.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; continue; }
            if (s.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; continue; }
            if (s.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; continue; }
            if (s.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; continue; }
        }
        o += s[i];
    }
    return o;
}

void XmlSerializer::serializeNode(const SerializationNode &node, std::ostream &out, int depth) {
    out << std::string(depth, '\t') << '<' << node.getName();
    for (const auto &kv : node.getProperties()) out << ' ' << kv.first << "=\"" << escape(kv.second) << '"';
    if (node.getChildren().empty()) {
        out << "/>\n";
        return;
    }
    out << ">\n";
    for (const auto &c : node.getChildren()) serializeNode(c, out, depth + 1);
    out << std::string(depth, '\t') << "</" << node.getName() << ">\n";
}

namespace {
struct Parser {
    std::string s;
    size_t i = 0;
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    [[noreturn]] void fail(const std::string &why) { throw OpenMMException("XML parse error: " + why); }
    std::string ident() {
        size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == ':' || s[i] == '-' || s[i] == '.')) i++;
        if (b == i) fail("name expected");
        return s.substr(b, i - b);
    }
    void element(SerializationNode &node) {
        if (s[i] != '<') fail("'<' expected");
        i++;
        node.setName(ident());
        for (;;) {
            ws();
            if (i >= s.size()) fail("unterminated element");
            if (s[i] == '/') {
                if (s.compare(i, 2, "/>") != 0) fail("'/>' expected");
                i += 2;
                return;
            }
            if (s[i] == '>') { i++; break; }
            std::string key = ident();
            ws();
            if (s[i] != '=') fail("'=' expected");
            i++;
            ws();
            char q = s[i];
            if (q != '"' && q != '\'') fail("quote expected");
            size_t e = s.find(q, i + 1);
            if (e == std::string::npos) fail("unterminated attribute");
            node.setStringProperty(key, unescape(s.substr(i + 1, e - i - 1)));
            i = e + 1;
        }
        for (;;) {
            ws();
            if (i >= s.size()) fail("unterminated element body");
            if (s.compare(i, 2, "</") == 0) {
                i += 2;
                if (ident() != node.getName()) fail("mismatched closing tag");
                ws();
                if (s[i] != '>') fail("'>' expected");
                i++;
                return;
            }
            if (s.compare(i, 4, "<!--") == 0) {
                size_t e = s.find("-->", i);
                if (e == std::string::npos) fail("unterminated comment");
                i = e + 3;
                continue;
            }
            if (s[i] != '<') { i++; continue; }  // ignore text content
            element(node.createChildNode(""));
        }
    }
};
}  // namespace

SerializationNode XmlSerializer::parse(std::istream &in) {
    Parser p;
    p.s.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    p.ws();
    if (p.s.compare(p.i, 5, "<?xml") == 0) {
        size_t e = p.s.find("?>", p.i);
        if (e == std::string::npos) p.fail("unterminated declaration");
        p.i = e + 2;
    }
    p.ws();
    SerializationNode root;
    p.element(root);
    return root;
}

}  // namespace OpenMM
#endif
