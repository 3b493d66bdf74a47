// CompiledLayers: the product of LayerFactory::compileLayers -- layers addressable by number and by name,
// iterated in ascending layer-number order (= execution order).  Reference: base/compiledlayers.h:58-286.
// Copies share ownership; the layer objects die with the last copy (:181-190).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <unordered_map>

#include "layerbase.h"

namespace fyusion {
namespace fyusenet {

class CompiledLayers {
    struct Store {
        std::map<int, LayerBase *> byNumber;
        std::unordered_map<std::string, LayerBase *> byName;
        ~Store() {
            for (auto &kv : byNumber) delete kv.second;
        }
    };

 public:
    // iterator exposing `.first` (layer number) and `.second` (layer) like the reference's iterator
    struct iterator {
        using inner = std::map<int, LayerBase *>::const_iterator;
        inner it;
        int first = -1;
        LayerBase *second = nullptr;
        explicit iterator(inner i, inner end) : it(i) { sync(end); endIt = end; }
        iterator &operator++() { ++it; sync(endIt); return *this; }
        bool operator!=(const iterator &o) const { return it != o.it; }
        bool operator==(const iterator &o) const { return it == o.it; }
        const iterator &operator*() const { return *this; }
     private:
        inner endIt;
        void sync(inner end) {
            if (it != end) { first = it->first; second = it->second; }
            else { first = -1; second = nullptr; }
        }
    };

    CompiledLayers() : store_(std::make_shared<Store>()) {}
    void setLayer(LayerBase *layer) {
        if (!layer) return;
        int no = layer->getNumber();
        if (no < 0) THROW_EXCEPTION_ARGS(FynException, "Illegal layer number %d for layer %s", no, layer->getName().c_str());
        if (store_->byNumber.count(no)) THROW_EXCEPTION_ARGS(FynException, "Layer number %d used twice (%s)", no, layer->getName().c_str());
        store_->byNumber[no] = layer;
        store_->byName[layer->getName()] = layer;
    }
    LayerBase *operator[](int number) const {
        auto it = store_->byNumber.find(number);
        return it == store_->byNumber.end() ? nullptr : it->second;
    }
    LayerBase *operator[](const std::string &name) const {
        auto it = store_->byName.find(name);
        return it == store_->byName.end() ? nullptr : it->second;
    }
    iterator begin() const { return iterator(store_->byNumber.begin(), store_->byNumber.end()); }
    iterator end() const { return iterator(store_->byNumber.end(), store_->byNumber.end()); }
    size_t size() const { return store_->byNumber.size(); }
    int maxLayerNumber() const { return store_->byNumber.empty() ? -1 : store_->byNumber.rbegin()->first; }
    void cleanup() {
        for (auto &kv : store_->byNumber) kv.second->cleanup();
    }

 private:
    std::shared_ptr<Store> store_;
};

}  // namespace fyusenet
}  // namespace fyusion
