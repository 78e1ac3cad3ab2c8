// openmm_standin.h -- the handful of OpenMM types the ATMMetaForce facade touches, for builds WITHOUT OpenMM
// (OpenMM is not installable in this image).  With -DATM_HAVE_OPENMM the facade includes the real headers instead
// and this file is not used.  Behaviour follows OpenMM's documented API (Force::getForceGroup/setForceGroup/getName/
// setName, OpenMMException, SerializationNode / SerializationProxy / XmlSerializer); nothing here is copied from OpenMM.
#ifndef ATM_OPENMM_STANDIN_H_
#define ATM_OPENMM_STANDIN_H_

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <istream>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace OpenMM {

class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string &message) : message(message) {}
    ~OpenMMException() throw() {}
    const char *what() const throw() { return message.c_str(); }

private:
    std::string message;
};

#define ASSERT_VALID_INDEX(index, vector)                                                    \
    {                                                                                        \
        if ((index) < 0 || (index) >= (long long)(vector).size())                            \
            throw OpenMM::OpenMMException("Assertion failure: Index out of range");         \
    }

class ForceImpl;
class Context;

class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) {
        if (group < 0 || group > 31) throw OpenMMException("Force group must be between 0 and 31");
        forceGroup = group;
    }
    const std::string &getName() const { return name; }
    void setName(const std::string &n) { name = n; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
    /** The per-Context implementation object, or NULL for a Force that is only data for another Force's Impl (here:
     *  NonbondedForce, evaluated by ATMMetaForceImpl).  OpenMM declares this protected and lets Context call it as a
     *  friend; the stand-in keeps it public. */
    virtual ForceImpl *createImpl() const { return nullptr; }
    /** A deep copy (what OpenMM's XmlSerializer::clone<Force> gives ATMMetaForceImpl::copysystem); NULL = not clonable. */
    virtual Force *clone() const { return nullptr; }

private:
    int forceGroup;
    std::string name;
};

// A tree of named nodes with string properties; numbers are stored in their shortest round-trip text form.
class SerializationNode {
public:
    const std::string &getName() const { return name; }
    void setName(const std::string &n) { name = n; }
    const std::vector<SerializationNode> &getChildren() const { return children; }
    const SerializationNode &getChildNode(const std::string &n) const {
        for (const auto &c : children)
            if (c.name == n) return c;
        throw OpenMMException("Unknown child node '" + n + "'");
    }
    SerializationNode &createChildNode(const std::string &n) {
        children.push_back(SerializationNode());
        children.back().name = n;
        return children.back();
    }
    const std::map<std::string, std::string> &getProperties() const { return properties; }
    bool hasProperty(const std::string &n) const { return properties.count(n) != 0; }
    const std::string &getStringProperty(const std::string &n) const {
        auto it = properties.find(n);
        if (it == properties.end()) throw OpenMMException("Unknown property '" + n + "' in node '" + name + "'");
        return it->second;
    }
    const std::string &getStringProperty(const std::string &n, const std::string &def) const {
        auto it = properties.find(n);
        return it == properties.end() ? def : it->second;
    }
    SerializationNode &setStringProperty(const std::string &n, const std::string &v) {
        properties[n] = v;
        return *this;
    }
    int getIntProperty(const std::string &n) const { return std::atoi(getStringProperty(n).c_str()); }
    int getIntProperty(const std::string &n, int def) const { return hasProperty(n) ? getIntProperty(n) : def; }
    SerializationNode &setIntProperty(const std::string &n, int v) { return setStringProperty(n, std::to_string(v)); }
    double getDoubleProperty(const std::string &n) const { return std::strtod(getStringProperty(n).c_str(), nullptr); }
    double getDoubleProperty(const std::string &n, double def) const { return hasProperty(n) ? getDoubleProperty(n) : def; }
    SerializationNode &setDoubleProperty(const std::string &n, double v) {
        char buf[64];
        int prec = 1;
        for (; prec <= 17; prec++) {  // shortest text that reads back to the same double
            std::snprintf(buf, sizeof(buf), "%.*g", prec, v);
            if (std::strtod(buf, nullptr) == v) break;
        }
        std::string text(buf);
        if (text.find('e') != std::string::npos && v == v) {  // plain decimal notation for ordinary magnitudes
            const double a = v < 0 ? -v : v;
            if (a >= 1e-5 && a < 1e15) {
                for (int dec = 0; dec <= 22; dec++) {
                    std::snprintf(buf, sizeof(buf), "%.*f", dec, v);
                    if (std::strtod(buf, nullptr) == v) break;
                }
                text = buf;
            }
        }
        return setStringProperty(n, text);
    }

private:
    std::string name;
    std::vector<SerializationNode> children;
    std::map<std::string, std::string> properties;
};

class SerializationProxy {
public:
    explicit SerializationProxy(const std::string &typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string &getTypeName() const { return typeName; }
    virtual void serialize(const void *object, SerializationNode &node) const = 0;
    virtual void *deserialize(const SerializationNode &node) const = 0;
    // registry keyed by type name (OpenMM keys it by typeid as well; one key is enough here)
    static std::map<std::string, const SerializationProxy *> &registry() {
        static std::map<std::string, const SerializationProxy *> r;
        return r;
    }
    static void registerProxy(const SerializationProxy *proxy) { registry()[proxy->getTypeName()] = proxy; }
    static const SerializationProxy &getProxy(const std::string &typeName) {
        auto it = registry().find(typeName);
        if (it == registry().end()) throw OpenMMException("No SerializationProxy registered for type '" + typeName + "'");
        return *it->second;
    }

private:
    std::string typeName;
};

// OpenMM's XML dialect: <?xml version="1.0" ?> then nested elements whose attributes are the node properties;
// the root element carries type="<proxy type name>".
class XmlSerializer {
public:
    static void serializeNode(const SerializationNode &node, std::ostream &out, int depth = 0);
    static SerializationNode parse(std::istream &in);

    template <class T>
    static void serialize(const T *object, const std::string &rootName, std::ostream &out, const std::string &typeName) {
        SerializationNode root;
        root.setName(rootName);
        const SerializationProxy &proxy = SerializationProxy::getProxy(typeName);
        proxy.serialize(object, root);
        root.setStringProperty("type", typeName);
        out << "<?xml version=\"1.0\" ?>\n";
        serializeNode(root, out);
    }
    template <class T>
    static T *deserialize(std::istream &in) {
        SerializationNode root = parse(in);
        const SerializationProxy &proxy = SerializationProxy::getProxy(root.getStringProperty("type"));
        return reinterpret_cast<T *>(proxy.deserialize(root));
    }
};

}  // namespace OpenMM

#endif
